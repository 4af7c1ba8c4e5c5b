"""Native executors of the generator / discriminator graphs: every layer is a call into libfpg_b200.so.

A network executor owns no parameters: it reads them from the drop-in nn.Module (models/model_architectures.py),
keeps bf16 GEMM-layout copies (repacked when the fp32 parameters change) and runs

    forward(x)            -> output, tape     (tape = saved activations of this call)
    backward(tape, grads) -> parameter gradients (+ input gradient)

Layer graphs follow the reference: model_architectures.py:339-400 (attention generators), :95-134 (CycleGAN
generator), :424-441 (InstanceNorm PatchGAN). Convolution biases that feed an InstanceNorm are mathematical
no-ops (the norm subtracts the per-plane mean) and are skipped; their gradient is exactly 0.
"""
import ctypes as C
import functools
import os

import torch

from . import lib as L
from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH, ActBuf, ConvSpec, pad16


# Precision of the tensors that no tensor-core instruction reads (both default on; FPG_PRENORM_F16=0 / FPG_SKIP_F16=0
# restore the all-bf16 storage of round 1 for A/B measurements)
PRENORM_F16 = os.environ.get("FPG_PRENORM_F16", "1") != "0"
SKIP_F16 = os.environ.get("FPG_SKIP_F16", "1") != "0"


class ConvLayer:
    """One convolution in the forward-conv view. For nn.ConvTranspose2d(Cin_T, Cout_T) the equivalent forward conv
    has c_out = Cin_T, c_in = Cout_T and the parameter [Cin_T][Cout_T][R][S] already is [K][C][R][S]."""

    def __init__(self, name, weight, bias, r, stride, pad, transposed=False, use_bias=False):
        self.name, self.weight, self.bias = name, weight, bias
        self.transposed, self.use_bias = transposed, use_bias
        k, c = weight.shape[0], weight.shape[1]
        self.k_valid, self.c_valid = k, c
        self.spec = ConvSpec(r, r, stride, pad, pad16(c), pad16(k), c_in_valid=c, c_out_valid=k)
        self.bias_pad = None
        self.s2d_spec = None  # _NetBase._add_s2d_stem: the same layer as a 2x2 stride-1 conv over a space-to-depth input

    def version(self):
        return (self.weight._version, self.weight.data_ptr(),
                self.bias._version if self.use_bias else 0, self.bias.data_ptr() if self.use_bias else 0)


def _s2d_job(layer, fprop_job):
    """pack job of the space-to-depth operand of a 4x4 stride-2 stem: [k][(ty, tx)][(i, j, c16)] bf16 -- the ordinary
    fprop operand [k][r * 4 + s][c16] with its taps permuted (r = 2 ty + i, s = 2 tx + j)"""
    sp = layer.s2d_spec
    if sp.w_fprop is None:
        nbytes = L.load().fpg_packed_weight_bytes(sp.gref())
        assert nbytes == 2 * fprop_job.rows * fprop_job.taps * fprop_job.cols and fprop_job.taps == 16
        sp.w_fprop = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device=layer.weight.device)
    job = _copy_job(fprop_job)
    job.dst = sp.w_fprop.data_ptr()
    for t in range(16):
        ty, tx, i, j = t >> 3, (t >> 2) & 1, (t >> 1) & 1, t & 1
        job.src_tap[t] = (2 * ty + i) * 4 + 2 * tx + j
    return job


class _PackTable:
    """Device-resident job table of fpg_pack_weights_batched for one network: every packed bf16 operand (fprop and
    dgrad layouts) and zero-padded bias vector is rebuilt from the fp32 parameters by ONE launch."""

    def __init__(self, layers):
        lib = L.load()
        jobs = []
        for layer in layers:
            w = layer.weight.detach()
            assert w.is_contiguous() and w.dtype == torch.float32 and w.is_cuda
            sp = layer.spec
            sp.alloc(w.device)
            rs = sp.g.r * sp.g.s
            buf = (L.PackJob * 5)()
            n = C.c_int32()
            L.call("fpg_pack_jobs", C.c_void_p(w.data_ptr()), sp.c_in_valid * rs, rs, sp.c_out_valid, sp.c_in_valid,
                   sp.gref(), C.c_void_p(sp.w_fprop.data_ptr()), C.c_void_p(sp.w_dgrad.data_ptr()), buf, C.byref(n))
            jobs.extend(_copy_job(buf[i]) for i in range(n.value))
            if layer.s2d_spec is not None:
                jobs.append(_s2d_job(layer, buf[0]))
            if layer.use_bias:
                npad = sp.g.c_in if layer.transposed else sp.g.c_out
                if layer.bias_pad is None:
                    layer.bias_pad = torch.zeros(npad, dtype=torch.float32, device=w.device)
                job = L.PackJob()
                L.call("fpg_pack_job_copy_f32", C.c_void_p(layer.bias.data_ptr()), layer.bias.numel(),
                       C.c_void_p(layer.bias_pad.data_ptr()), npad, C.byref(job))
                jobs.append(job)
        arr = (L.PackJob * len(jobs))(*jobs)
        block_job, block_first = [], []
        for j, job in enumerate(jobs):
            nb = lib.fpg_pack_job_blocks(C.byref(job))
            block_job.extend([j] * nb)
            block_first.extend(range(nb))
        dev = layers[0].weight.device
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.jobs = raw.to(dev)
        self.block_job = torch.tensor(block_job, dtype=torch.int32, device=dev)
        self.block_first = torch.tensor(block_first, dtype=torch.int32, device=dev)
        self.n_blocks = len(block_job)

    def run(self):
        ops.pack_weights_batched(self.jobs, self.block_job, self.block_first, self.n_blocks)


class AdamPack:
    """Adam + the repack of every bf16 GEMM operand of the networks of one optimiser in ONE launch
    (fpg_adam_pack_step, csrc/adam_pack.cu): convolution weights are updated tile by tile and written straight into their
    fprop / data-gradient operand layouts from shared memory; every other parameter (and the zero-padded bias vectors the
    epilogues read) by chunk blocks of the same launch. `fp` is a trainer.FlatParams over the modules that `executors`
    run; the operands are packed once here the old way so that their padding is in place."""

    CHUNK = 4096

    def __init__(self, fp, executors):
        lib = L.load()
        self.fp = fp
        for ex in executors:
            ex.repack(force=True)
        base, total = fp.flat.data_ptr(), fp.flat.numel()
        dev = fp.flat.device
        jobs, layers, covered, biases = [], [], [], []
        for ex in executors:
            for layer in ex.layers.values():
                w = layer.weight
                off = (w.data_ptr() - base) // 4
                assert 0 <= off and off + w.numel() <= total, "the executor's parameters do not live in this flat buffer"
                sp = layer.spec
                rs = sp.g.r * sp.g.s
                assert w.numel() == sp.c_out_valid * sp.c_in_valid * rs
                buf = (L.PackJob * 5)()
                n = C.c_int32()
                L.call("fpg_pack_jobs", C.c_void_p(w.data_ptr()), sp.c_in_valid * rs, rs, sp.c_out_valid, sp.c_in_valid,
                       sp.gref(), C.c_void_p(sp.w_fprop.data_ptr()), C.c_void_p(sp.w_dgrad.data_ptr()), buf, C.byref(n))
                tk, tc = C.c_int32(), C.c_int32()
                L.call("fpg_adam_pack_tile", rs, C.byref(tk), C.byref(tc))
                lay = L.AdamPackLayer()
                lay.p_off, lay.k, lay.c, lay.rs = off, sp.c_out_valid, sp.c_in_valid, rs
                lay.tk, lay.tc, lay.n_jobs = tk.value, tc.value, n.value
                for i in range(n.value):
                    lay.job[i] = len(jobs)
                    jobs.append(_copy_job(buf[i]))
                if layer.s2d_spec is not None:
                    lay.job[lay.n_jobs] = len(jobs)
                    lay.n_jobs += 1
                    jobs.append(_s2d_job(layer, buf[0]))
                layers.append(lay)
                covered.append((off, off + w.numel()))
                if layer.use_bias:
                    boff = (layer.bias.data_ptr() - base) // 4
                    assert 0 <= boff and boff + layer.bias.numel() <= total and layer.bias_pad is not None
                    biases.append((boff, boff + layer.bias.numel(), layer.bias_pad.data_ptr()))
        chunks = []
        for a, b in self._gaps(sorted(covered), total):
            cuts = sorted({a, b} | {x for lo, hi, _ in biases for x in (lo, hi) if a < x < b})
            for x, y in zip(cuts, cuts[1:]):
                dst = next((ptr + 4 * (x - lo) for lo, hi, ptr in biases if lo <= x and y <= hi), None)
                for o in range(x, y, self.CHUNK):
                    ch = L.AdamPackChunk()
                    ch.off, ch.count = o, min(self.CHUNK, y - o)
                    ch.copy_dst = dst + 4 * (o - x) if dst is not None else None
                    chunks.append(ch)
        item, first = [], []
        for i, lay in enumerate(layers):
            nt = -(-lay.k // lay.tk) * -(-lay.c // lay.tc)
            item.extend([i] * nt)
            first.extend(range(nt))
        for i in range(len(chunks)):
            item.append(-1 - i)
            first.append(0)
        self.n_params_tiled = sum(b - a for a, b in covered)

        def to_dev(arr_type, items):
            if not items:
                return torch.zeros(8, dtype=torch.uint8, device=dev)
            arr = (arr_type * len(items))(*items)
            return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)

        self.layers = to_dev(L.AdamPackLayer, layers)
        self.jobs = to_dev(L.PackJob, jobs)
        self.chunks = to_dev(L.AdamPackChunk, chunks)
        self.block_item = torch.tensor(item, dtype=torch.int32, device=dev)
        self.block_first = torch.tensor(first, dtype=torch.int32, device=dev)
        self.n_blocks = len(item)
        self._own = (C.c_void_p * 1)(fp.grads.flat.data_ptr())

    @staticmethod
    def _gaps(covered, total):
        pos = 0
        for a, b in covered:
            assert a >= pos, "overlapping parameters"
            if a > pos:
                yield pos, a
            pos = b
        if pos < total:
            yield pos, total

    def run(self, grad_scale=1.0, sources=None, n_src=1, gsum=None, betas=(0.5, 0.999), eps=1e-8):
        """one optimiser step on the gradient fp.grads.flat (or the rank-ordered sum of `sources`, a ctypes array of
        n_src device pointers), operands repacked"""
        fp = self.fp
        ops._run("adam_step", 2, "fpg_adam_pack_step", ops._ptr(fp.flat), sources if sources is not None else self._own,
                 n_src, ops._ptr(fp.m), ops._ptr(fp.v), float(betas[0]), float(betas[1]), float(eps), ops._ptr(fp.state),
                 float(grad_scale), ops._ptr(gsum) if gsum is not None else None, ops._ptr(self.layers),
                 ops._ptr(self.jobs), ops._ptr(self.chunks), ops._ptr(self.block_item), ops._ptr(self.block_first),
                 self.n_blocks, ops._stream())


def _copy_job(job):
    out = L.PackJob()
    C.memmove(C.byref(out), C.byref(job), C.sizeof(L.PackJob))
    return out


class Grads:
    """Destination of parameter gradients: name -> fp32 tensor shaped like the parameter (views of a flat buffer)."""

    def __init__(self, named_params, flat=None):
        self.names = [n for n, _ in named_params]
        sizes = [p.numel() for _, p in named_params]
        total = sum(sizes)
        dev = named_params[0][1].device
        self.flat = flat if flat is not None else torch.zeros(total, dtype=torch.float32, device=dev)
        self.views = {}
        off = 0
        for (n, p), s in zip(named_params, sizes):
            self.views[n] = self.flat[off:off + s].view(p.shape)
            off += s

    def __getitem__(self, name):
        return self.views[name]


def _norm_act(y, act, halo, residual=None):
    """IN statistics + apply; returns (stats, z)"""
    stats = torch.empty(y.n * y.c * 2, dtype=torch.float32, device=y.t.device)
    ops.instnorm_stats(y, stats)
    z = ActBuf(y.n, y.h, y.w, y.c, halo=halo, zero=False)
    ops.instnorm_apply(y, stats, act, z, residual=residual)
    return stats, z


class _NetBase:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fn = cls.__dict__.get("backward")
        if fn is not None:  # every backward() joins the weight-gradient stream before it returns
            @functools.wraps(fn)
            def backward(self, *a, **k):
                try:
                    return fn(self, *a, **k)
                finally:
                    ops.wgrad_join()
            cls.backward = backward

    def __init__(self, module):
        self.module = module
        self.layers = {}
        self._pack = None
        self._pack_ptrs = None
        self._pack_ver = None
        self.grad_ready = None  # optional callback(param_name) fired once a parameter gradient has been produced

    def _add(self, name, conv_module, r, stride, pad, transposed=False, use_bias=False):
        layer = ConvLayer(name, conv_module.weight, conv_module.bias, r, stride, pad, transposed, use_bias)
        self.layers[name] = layer
        return layer

    def _add_s2d_stem(self, name):
        """EXPERIMENT (FPG_S2D_STEM=1, off by default): the 4x4 stride-2 pad-1 layer `name` over a 16-channel input (the
        PatchGAN stem) ALSO as a 2x2 stride-1 convolution over the space-to-depth copy of its input
        (ops.space_to_depth16): 128-byte TMA rows and 4 taps instead of 32-byte rows and 16 taps. Forward only; the
        gradients keep the ordinary formulation. Parity-green, but measured NOT faster (tools/micro_s2d.py, B=32:
        strided 102 us; copy 27 us + 2x2 kernel 98 us tiled / 145 us row-stationary): the layer is bound by the
        L2 -> SM path (402 MB through the crossbar for a 67 MB input: every tile re-fetches the 32 KB of weights and
        every input pixel is fetched by 4 taps, ncu r02_m0_fprop), not by the TMA row rate, and the space-to-depth form
        moves the same bytes."""
        layer = self.layers[name]
        g = layer.spec.g
        assert (g.r, g.s, g.stride, g.pad, g.c_in) == (4, 4, 2, 1, 16) and not layer.transposed
        if os.environ.get("FPG_S2D_STEM", "0") == "1":
            layer.s2d_spec = ConvSpec(2, 2, 1, 0, 64, g.c_out, c_in_valid=64, c_out_valid=layer.spec.c_out_valid)

    def _stem_conv(self, din, name, act, xs=None):
        """forward of a stem registered with _add_s2d_stem -> (output, space-to-depth copy of din or None); `xs`: that
        copy when the caller already has it (the same discriminator input is used twice per training step)"""
        layer = self.layers[name]
        if layer.s2d_spec is None or din.h % 2 or din.w % 2 or din.halo or din.c_stride != 16:
            return self._conv(din, name, act=act), None
        if xs is None:
            xs = ActBuf(din.n, din.h // 2 + 1, din.w // 2 + 1, 64, zero=False)
            ops.space_to_depth16(din, xs)
        y = ActBuf(din.n, din.h // 2, din.w // 2, layer.spec.g.c_out, zero=False)
        ops.conv_fprop(xs, layer.s2d_spec, y, bias=layer.bias_pad if layer.use_bias else None, act=act)
        return y, xs

    def repack(self, force=False):
        """Rebuild the bf16 GEMM operands if any fp32 parameter changed since the last call (one launch)."""
        ver = tuple(layer.version() for layer in self.layers.values())
        ptrs = tuple(v[1::2] for v in ver)
        if self._pack is None or ptrs != self._pack_ptrs:
            self._pack = _PackTable(list(self.layers.values()))  # parameters were re-homed (.cuda(), FlatParams)
            self._pack_ptrs = ptrs
        elif ver == self._pack_ver and not force:
            return
        self._pack_ver = ver
        self._pack.run()

    def named_params(self):
        return list(self.module.named_parameters())

    # conv helpers ------------------------------------------------------------------------------
    def _conv(self, x, name, halo_out=0, fp32=False, act=ACT_NONE):
        """forward conv (or transposed-conv forward) of layer `name` -> raw output buffer"""
        L = self.layers[name]
        g = L.spec.g
        bias = L.bias_pad if L.use_bias else None
        if L.transposed:
            y = ActBuf(x.n, x.h * 2, x.w * 2, g.c_in, halo=halo_out, fp32=fp32, zero=False)
            ops.conv_dgrad(x, L.spec, y, bias=bias, act=act)
        else:
            hp, wp = x.h + 2 * x.halo, x.w + 2 * x.halo
            ho = (hp + 2 * g.pad - g.r) // g.stride + 1
            wo = (wp + 2 * g.pad - g.s) // g.stride + 1
            y = ActBuf(x.n, ho, wo, g.c_out, halo=halo_out, fp32=fp32, zero=False)
            ops.conv_fprop(x, L.spec, y, bias=bias, act=act)
        return y

    def _conv_stats(self, x, name, batch=False, eps=1e-5):
        """Raw forward conv of a normalised layer plus its {mean, rstd}: from the conv epilogue where the kernel has
        that epilogue (no separate pass over the activation), else from the statistics kernel.
        The pre-InstanceNorm output is stored as fp16 (PRENORM_F16): only elementwise kernels read it, so the
        conv-output rounding site of the forward drops from 2^-9 to 2^-12 relative at the same traffic
        (tests/sim_bf16_rounding.py: generator output 2.27e-2 -> 1.84e-2 away from fp32)."""
        L = self.layers[name]
        g = L.spec.g
        f16 = PRENORM_F16 and not batch
        if L.transposed:
            y = ActBuf(x.n, x.h * 2, x.w * 2, g.c_in, zero=False, f16=f16)
        else:
            hp, wp = x.h + 2 * x.halo, x.w + 2 * x.halo
            y = ActBuf(x.n, (hp + 2 * g.pad - g.r) // g.stride + 1, (wp + 2 * g.pad - g.s) // g.stride + 1, g.c_out,
                       zero=False, f16=f16)
        stats = torch.empty((1 if batch else y.n) * y.c * 2, dtype=torch.float32, device=y.t.device)
        if not ops.conv_with_stats(x, L.spec, y, stats, transposed=L.transposed, eps=eps, batch=batch):
            if L.transposed:
                ops.conv_dgrad(x, L.spec, y)
            else:
                ops.conv_fprop(x, L.spec, y)
            if batch:
                ops.batch_stats(y, stats, eps)
            else:
                ops.instnorm_stats(y, stats, eps)
        return y, stats

    def _conv_in(self, x, name, act, halo, residual=None, skip_out=None):
        """conv -> InstanceNorm -> activation (+ residual, + reflect halo); returns (y, stats, z)"""
        y, stats = self._conv_stats(x, name)
        z = ActBuf(y.n, y.h, y.w, y.c, halo=halo, zero=False)
        ops.instnorm_apply(y, stats, act, z, residual=residual, skip_out=skip_out)
        return y, stats, z

    def _conv_bwd(self, name, x, dy, grads, need_dx, dx_halo=0, in_bwd=None):
        """Backward of layer `name`: x = saved layer input, dy = gradient of its raw output.
        Writes the weight (and used bias) gradient into `grads`; returns dx (incl. halo when dx_halo > 0).
        in_bwd = (xprev, add): x is the output of an InstanceNorm block -- relu(IN(y)) when xprev is None, else
        xprev + IN(y) -- and the data-gradient kernel also produces the plane means of that InstanceNorm's backward
        from x (and xprev), merging the skip gradient `add` into dx's interior; returns (dx, red), red = None when the
        layer has no such epilogue (dx is then the plain gradient, `add` NOT merged)."""
        L = self.layers[name]
        g = L.spec.g
        if grads is not None:
            gw = grads[name + ".weight"]
            with ops.wgrad_side(x, dy):  # second stream, joined at the end of backward()
                if L.transposed:
                    ops.conv_wgrad(dy, x, L.spec, gw)  # conv view: input role = dy_T (large), output-grad role = x_T
                else:
                    ops.conv_wgrad(x, dy, L.spec, gw)
                if L.use_bias:
                    ops.bias_grad(dy, grads[name + ".bias"], L.bias.numel())
                if self.grad_ready is not None:  # collectives are ordered after the stream they are launched from
                    self.grad_ready(name + ".weight")
                    if L.bias is not None:
                        self.grad_ready(name + ".bias")  # unused biases (before an InstanceNorm) keep gradient 0
        if not need_dx:
            return None
        if L.transposed:
            dx = ActBuf(x.n, x.h, x.w, g.c_out, halo=0, zero=False)
            ops.conv_fprop(dy, L.spec, dx)
        else:
            dx = ActBuf(x.n, x.h, x.w, g.c_in, halo=dx_halo, zero=False)
            if in_bwd is not None:  # fuse the reduction pass of the producing InstanceNorm's backward (ops.py)
                xprev, add = in_bwd
                red = ops.conv_dgrad_inbwd(dy, L.spec, dx, x, xprev, add)
                if red is not None:
                    return dx, red
            ops.conv_dgrad(dy, L.spec, dx)
            if in_bwd is not None:
                return dx, None
        return dx


class _ResnetGeneratorNet(_NetBase):
    """Shared trunk of the ResNet generators: reflect-pad 7x7 stem, two stride-2 convs, 9 residual blocks
    (model_architectures.py:342-346 + :412-418, :95-105 + :122-134), and the two-stage transposed-conv decoder."""

    # names of the trunk layers in the module's state_dict: overridden per architecture
    STEM = ("conv1", "conv2", "conv3")

    def _block_names(self, i):
        raise NotImplementedError

    def _trunk_forward(self, x, t, xin=None):
        """x: fp32 NCHW input, or xin: the input already packed (bf16 NHWC, 16 channels, reflect halo 3)"""
        c1, c2, c3 = self.STEM
        if xin is None:
            B, C, H, W = x.shape
            xin = ActBuf(B, H, W, 16, halo=3, zero=False)
            ops.pack_nchw(x, xin, 0, zero_rest=True)
        assert xin.halo == 3 and xin.c == 16
        t["xin"] = xin
        t["y1"], t["s1"], t["z1"] = self._conv_in(t["xin"], c1, ACT_RELU, 0)
        t["y2"], t["s2"], t["z2"] = self._conv_in(t["z1"], c2, ACT_RELU, 0)
        # the residual stream x_i is carried twice: bf16 with its reflect halo (the tensor-core operand of the next
        # convolution, saved for the weight gradient) and halo-free fp16 (the value the next skip addition reads), so
        # that the nine additions of model_architectures.py:412-418 do not re-round it to 8 mantissa bits each time
        def skip_buf(like):
            return ActBuf(like.n, like.h, like.w, like.c, zero=False, f16=True) if SKIP_F16 else None

        y3, s3 = self._conv_stats(t["z2"], c3)
        xcur = ActBuf(y3.n, y3.h, y3.w, y3.c, halo=1, zero=False)
        skip = skip_buf(y3)
        ops.instnorm_apply(y3, s3, ACT_RELU, xcur, skip_out=skip)
        t["y3"], t["s3"] = y3, s3
        t["x0"] = xcur
        for i in range(self.n_blocks):
            n1, n2 = self._block_names(i)
            ya, sa, za = self._conv_in(xcur, n1, ACT_RELU, 1)
            last = i == self.n_blocks - 1
            skip_next = None if last else skip_buf(ya)
            yb, sb, xnext = self._conv_in(za, n2, ACT_NONE, 0 if last else 1,
                                          residual=skip if skip is not None else xcur, skip_out=skip_next)
            t[f"b{i}"] = (ya, sa, za, yb, sb)
            t[f"x{i + 1}"] = xnext
            xcur, skip = xnext, skip_next
        return xcur

    def _decoder_forward(self, x, up1, up2, halo_out):
        u1, s1, v1 = self._conv_in(x, up1, ACT_RELU, 0)
        u2, s2, v2 = self._conv_in(v1, up2, ACT_RELU, halo_out)
        return (u1, s1, v1, u2, s2, v2)

    def _decoder_backward(self, saved, x_top, dv2, up1, up2, grads):
        """dv2: gradient w.r.t. the decoder output v2 (incl. halo). Returns the gradient w.r.t. the trunk output."""
        u1, s1, v1, u2, s2, v2 = saved
        du2 = ActBuf(u2.n, u2.h, u2.w, u2.c, zero=False)
        ops.instnorm_bwd(dv2, u2, s2, ACT_RELU, du2)
        dv1 = self._conv_bwd(up2, v1, du2, grads, True)
        du1 = ActBuf(u1.n, u1.h, u1.w, u1.c, zero=False)
        ops.instnorm_bwd(dv1, u1, s1, ACT_RELU, du1)
        return self._conv_bwd(up1, x_top, du1, grads, True)

    def _trunk_backward(self, t, dz, dz2, grads, need_dx):
        """dz (+ dz2): gradient w.r.t. the trunk output x_n (interior). Returns d(xin) incl. halo if need_dx."""
        c1, c2, c3 = self.STEM
        # `red` is not None: dz already holds the merged gradient fold^-1(dz) + dz2 w.r.t. x_{i+1} and `red` the plane
        # means of the InstanceNorm backward that consumes it -- both produced by the previous data-gradient kernel
        red = None
        for i in reversed(range(self.n_blocks)):
            ya, sa, za, yb, sb = t[f"b{i}"]
            n1, n2 = self._block_names(i)
            xi = t[f"x{i}"]
            dyb = ActBuf(yb.n, yb.h, yb.w, yb.c, zero=False)
            if red is not None:
                ops.instnorm_bwd_apply(dz, yb, sb, red, ACT_NONE, dyb)
                skip = dz  # folded in place: its interior is the total gradient w.r.t. x_{i+1}
            else:
                gres = None
                if dz2 is not None or dz.halo:
                    gres = ActBuf(yb.n, yb.h, yb.w, yb.c, zero=False)  # total gradient w.r.t. x_{i+1} (skip branch)
                ops.instnorm_bwd(dz, yb, sb, ACT_NONE, dyb, dz2=dz2, dres=gres)
                skip = gres if gres is not None else dz
            dza, red_a = self._conv_bwd(n2, za, dyb, grads, True, dx_halo=1, in_bwd=(None, None))  # za = relu(IN(ya))
            dya = ActBuf(ya.n, ya.h, ya.w, ya.c, zero=False)
            if red_a is not None:
                ops.instnorm_bwd_apply(dza, ya, sa, red_a, ACT_RELU, dya)
            else:
                ops.instnorm_bwd(dza, ya, sa, ACT_RELU, dya)
            # x_i = act(IN(y_prev)) + (residual input of block i-1): block i-1's second norm, or conv3's for block 0
            y_prev, s_prev, act_prev = (t[f"b{i - 1}"][3], t[f"b{i - 1}"][4], ACT_NONE) if i > 0 else \
                (t["y3"], t["s3"], ACT_RELU)
            # x_i = x_{i-1} + IN(yb of block i-1) for i > 0, relu(IN(y3)) for the trunk input
            dz, red = self._conv_bwd(n1, xi, dya, grads, True, dx_halo=1,
                                     in_bwd=(t[f"x{i - 1}"] if i > 0 else None, skip))
            dz2 = None if red is not None else skip
        dy3 = ActBuf(t["y3"].n, t["y3"].h, t["y3"].w, t["y3"].c, zero=False)
        if red is not None:
            ops.instnorm_bwd_apply(dz, t["y3"], t["s3"], red, ACT_RELU, dy3)
        else:
            ops.instnorm_bwd(dz, t["y3"], t["s3"], ACT_RELU, dy3, dz2=dz2)
        dz2_ = self._conv_bwd(c3, t["z2"], dy3, grads, True)
        dy2 = ActBuf(t["y2"].n, t["y2"].h, t["y2"].w, t["y2"].c, zero=False)
        ops.instnorm_bwd(dz2_, t["y2"], t["s2"], ACT_RELU, dy2)
        dz1 = self._conv_bwd(c2, t["z1"], dy2, grads, True)
        dy1 = ActBuf(t["y1"].n, t["y1"].h, t["y1"].w, t["y1"].c, zero=False)
        ops.instnorm_bwd(dz1, t["y1"], t["s1"], ACT_RELU, dy1)
        return self._conv_bwd(c1, t["xin"], dy1, grads, need_dx, dx_halo=3)

    def _input_grad_nchw(self, t, dxin, extra_rgb=None, rgb_only=False):
        """fp32 NCHW gradient w.r.t. the network input from the haloed NHWC gradient of the stem (+ extra_rgb, the
        gradient the blend sends to the pre-flood RGB directly). rgb_only: only the three image channels
        [B, 3, H, W] (what train_cycle's second generator pass sends back to the first), native kernels only."""
        xin = t["xin"]
        folded = ActBuf(xin.n, xin.h, xin.w, 16, zero=False)
        ops.halo_fold(dxin, None, folded)
        if rgb_only:
            if extra_rgb is not None:
                ops.unpack_nchw(folded, extra_rgb, 0, accumulate=True)
                return extra_rgb
            dx = torch.empty(xin.n, 3, xin.h, xin.w, dtype=torch.float32, device=xin.t.device)
            ops.unpack_nchw(folded, dx, 0)
            return dx
        c_in = self.layers[self.STEM[0]].c_valid
        dx = torch.zeros(xin.n, c_in, xin.h, xin.w, dtype=torch.float32, device=xin.t.device)
        ops.unpack_nchw(folded, dx, 0)
        if extra_rgb is not None:
            dx[:, :3] += extra_rgb
        return dx


class AttentionGeneratorNet(_ResnetGeneratorNet):
    """PairedAttentionGenerator / AttentionGANGenerator (model_architectures.py:305-400, 163-258)."""

    def __init__(self, module):
        super().__init__(module)
        m = module
        self._add("conv1", m.conv1, 7, 1, 0)
        self._add("conv2", m.conv2, 3, 2, 1)
        self._add("conv3", m.conv3, 3, 2, 1)
        for i, blk in enumerate(m.resnet_blocks):
            self._add(f"resnet_blocks.{i}.conv1", blk.conv1, 3, 1, 0)
            self._add(f"resnet_blocks.{i}.conv2", blk.conv2, 3, 1, 0)
        for br in ("content", "attention"):
            self._add(f"deconv1_{br}", getattr(m, f"deconv1_{br}"), 3, 2, 1, transposed=True)
            self._add(f"deconv2_{br}", getattr(m, f"deconv2_{br}"), 3, 2, 1, transposed=True)
        self._add("deconv3_content", m.deconv3_content, 7, 1, 0, use_bias=True)
        self._add("deconv3_attention", m.deconv3_attention, 1, 1, 0, use_bias=True)
        self.n_blocks = len(m.resnet_blocks)

    def _block_names(self, i):
        return f"resnet_blocks.{i}.conv1", f"resnet_blocks.{i}.conv2"

    def forward(self, x, d_input=None, d_c0=0, want_nchw=True, xin=None):
        """x: fp32 NCHW [B, C<=16, H, W] (or xin: the packed input, see _trunk_forward). Returns (out_nchw, tape). If
        d_input (ActBuf [B,H,W,16]) is given the generated image is also written as bf16 into its channels
        [d_c0, d_c0+3)."""
        self.repack()
        t = {}
        xtop = self._trunk_forward(x, t, xin)
        B, H, W = t["xin"].n, t["xin"].h, t["xin"].w
        dev = t["xin"].t.device
        t["content"] = self._decoder_forward(xtop, "deconv1_content", "deconv2_content", 3)
        t["attention"] = self._decoder_forward(xtop, "deconv1_attention", "deconv2_attention", 0)
        t["c"] = self._conv(t["content"][5], "deconv3_content", fp32=True, act=ACT_TANH)
        t["l"] = self._conv(t["attention"][5], "deconv3_attention", fp32=True)
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if want_nchw else None
        mask = torch.empty(B, H, W, dtype=torch.float32, device=dev)
        ops.blend_fwd(t["c"], t["l"], t["xin"], out=d_input, out_c0=d_c0, out_nchw=out, mask=mask)
        t["mask"] = mask
        return out, t

    def backward(self, t, grads, dout_nchw=None, dout_nhwc=None, dout_c0=0, need_dx=False, rgb_only=False):
        """Writes every parameter gradient of this call into `grads` (overwrite, no accumulation).
        Returns the fp32 NCHW input gradient if need_dx (rgb_only: its three image channels only)."""
        xin = t["xin"]
        B, H, W = xin.n, xin.h, xin.w
        dc = ActBuf(B, H, W, 32, zero=False)
        dl = ActBuf(B, H, W, 16, zero=False)
        dimg = torch.empty(B, 3, H, W, dtype=torch.float32, device=xin.t.device) if need_dx else None
        ops.blend_bwd(t["c"], t["l"], xin, dc, dl, dout_nchw=dout_nchw, dout_nhwc=dout_nhwc, dout_c0=dout_c0,
                      dimage_nchw=dimg)
        xtop = t[f"x{self.n_blocks}"]
        gx = []
        for br, dhead, head in (("content", dc, "deconv3_content"), ("attention", dl, "deconv3_attention")):
            v2 = t[br][5]
            dv2 = self._conv_bwd(head, v2, dhead, grads, True, dx_halo=v2.halo)
            gx.append(self._decoder_backward(t[br], xtop, dv2, f"deconv1_{br}", f"deconv2_{br}", grads))
        dxin = self._trunk_backward(t, gx[0], gx[1], grads, need_dx)  # x_n gradient = sum of both decoder branches
        return self._input_grad_nchw(t, dxin, dimg, rgb_only) if need_dx else None


class CycleGANGeneratorNet(_ResnetGeneratorNet):
    """CycleGANGenerator (model_architectures.py:91-134): same trunk, one decoder, 7x7 -> 3 + tanh head."""

    STEM = ("model.1", "model.4", "model.7")

    def __init__(self, module):
        super().__init__(module)
        seq = module.model
        self._add("model.1", seq[1], 7, 1, 0)
        self._add("model.4", seq[4], 3, 2, 1)
        self._add("model.7", seq[7], 3, 2, 1)
        self.n_blocks = 9
        for i in range(9):
            blk = seq[10 + i].conv_block
            self._add(f"model.{10 + i}.conv_block.1", blk[1], 3, 1, 0)
            self._add(f"model.{10 + i}.conv_block.5", blk[5], 3, 1, 0)
        self._add("model.19", seq[19], 3, 2, 1, transposed=True)
        self._add("model.22", seq[22], 3, 2, 1, transposed=True)
        self._add("model.26", seq[26], 7, 1, 0, use_bias=True)

    def _block_names(self, i):
        return f"model.{10 + i}.conv_block.1", f"model.{10 + i}.conv_block.5"

    def forward(self, x, d_input=None, d_c0=0, xin=None):
        """as AttentionGeneratorNet.forward"""
        self.repack()
        t = {}
        xtop = self._trunk_forward(x, t, xin)
        B, H, W = t["xin"].n, t["xin"].h, t["xin"].w
        t["dec"] = self._decoder_forward(xtop, "model.19", "model.22", 3)
        t["o"] = self._conv(t["dec"][5], "model.26", fp32=True, act=ACT_TANH)  # fp32 NHWC, 3 of 16 channels valid
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=t["xin"].t.device)
        ops.unpack_nchw(t["o"], out, 0)
        if d_input is not None:
            ops.pack_nchw(out, d_input, d_c0)
        return out, t

    def backward(self, t, grads, dout_nchw, need_dx=False, rgb_only=False):
        o = t["o"]
        dpre = ActBuf(o.n, o.h, o.w, 16, zero=False)
        ops.tanh_bwd_pack(dout_nchw, o, dpre)
        v2 = t["dec"][5]
        dv2 = self._conv_bwd("model.26", v2, dpre, grads, True, dx_halo=3)
        gtop = self._decoder_backward(t["dec"], t[f"x{self.n_blocks}"], dv2, "model.19", "model.22", grads)
        dxin = self._trunk_backward(t, gtop, None, grads, need_dx)
        return self._input_grad_nchw(t, dxin, rgb_only=rgb_only) if need_dx else None


class PatchGANNet(_NetBase):
    """InstanceNorm 70x70 PatchGAN (model_architectures.py:420-441, 136-157, 278-299)."""

    def __init__(self, module):
        super().__init__(module)
        seq = module.model
        self._add("model.0", seq[0], 4, 2, 1, use_bias=True)
        self._add("model.2", seq[2], 4, 2, 1)
        self._add("model.5", seq[5], 4, 2, 1)
        self._add("model.8", seq[8], 4, 1, 1)
        self._add("model.11", seq[11], 4, 1, 1, use_bias=True)
        self._add_s2d_stem("model.0")

    def forward_buf(self, din, xs=None):
        """din: ActBuf [B, H, W, 16] (D input, channels beyond the real ones zero); xs: its space-to-depth copy if the
        caller kept it from an earlier pass (tape["din_s2d"]). Returns (logits ActBuf fp32, tape)."""
        self.repack()
        t = {"din": din}
        t["a1"], t["din_s2d"] = self._stem_conv(din, "model.0", ACT_LEAKY, xs)
        t["y2"], t["s2"], t["a2"] = self._conv_in(t["a1"], "model.2", ACT_LEAKY, 0)
        t["y3"], t["s3"], t["a3"] = self._conv_in(t["a2"], "model.5", ACT_LEAKY, 0)
        t["y4"], t["s4"], t["a4"] = self._conv_in(t["a3"], "model.8", ACT_LEAKY, 0)
        t["logits"] = self._conv(t["a4"], "model.11", fp32=True)
        return t["logits"], t

    def forward(self, x):
        """x: fp32 NCHW [B, C<=16, H, W] -> (logits fp32 NCHW [B,1,h,w], tape)"""
        B, C, H, W = x.shape
        din = ActBuf(B, H, W, 16, zero=False)
        ops.pack_nchw(x, din, 0, zero_rest=True)
        logits, t = self.forward_buf(din)
        return logits.to_nchw(1), t

    def backward(self, t, dlogits, grads, need_dx):
        """dlogits: ActBuf bf16 [B,h,w,16] (channel 0 valid). grads may be None (parameters frozen).
        Returns d(din) ActBuf [B,H,W,16] if need_dx."""
        da4 = self._conv_bwd("model.11", t["a4"], dlogits, grads, True)
        dy4 = ActBuf(t["y4"].n, t["y4"].h, t["y4"].w, t["y4"].c, zero=False)
        ops.instnorm_bwd(da4, t["y4"], t["s4"], ACT_LEAKY, dy4)
        da3 = self._conv_bwd("model.8", t["a3"], dy4, grads, True)
        dy3 = ActBuf(t["y3"].n, t["y3"].h, t["y3"].w, t["y3"].c, zero=False)
        ops.instnorm_bwd(da3, t["y3"], t["s3"], ACT_LEAKY, dy3)
        da2 = self._conv_bwd("model.5", t["a2"], dy3, grads, True)
        dy2 = ActBuf(t["y2"].n, t["y2"].h, t["y2"].w, t["y2"].c, zero=False)
        ops.instnorm_bwd(da2, t["y2"], t["s2"], ACT_LEAKY, dy2)
        da1 = self._conv_bwd("model.2", t["a1"], dy2, grads, True)
        dy1 = ActBuf(t["a1"].n, t["a1"].h, t["a1"].w, t["a1"].c, zero=False)
        ops.act_bwd(da1, t["a1"], ACT_LEAKY, dy1)
        return self._conv_bwd("model.0", t["din"], dy1, grads, need_dx)


# ------------------------------------------------------------------------------------------------ Pix2Pix
class _BatchNormMixin:
    """BatchNorm2d in training mode for the executors below: batch statistics + running-statistics side effects."""

    def _bn_forward(self, y, bn, act1, z1, act2=ACT_NONE, z2=None, mask=None, stats=None):
        if stats is None:
            stats = torch.empty(y.c * 2, dtype=torch.float32, device=y.t.device)
            ops.batch_stats(y, stats, bn.eps)
        if bn.track_running_stats and bn.training:
            ops.batchnorm_running_update(stats, y.n * y.h * y.w, bn.running_mean, bn.running_var, bn.eps,
                                         bn.momentum)
            bn.num_batches_tracked += 1
        ops.batchnorm_apply(y, stats, bn.weight.detach(), bn.bias.detach(), act1, z1, act2, z2, mask=mask)
        return stats

    def _bn_backward(self, name, bn, dz1, act1, y, stats, grads, dz2=None, act2=ACT_NONE, mask=None):
        dy = ActBuf(y.n, y.h, y.w, y.c, zero=False)
        dg = grads[name + ".weight"] if grads is not None else None
        db = grads[name + ".bias"] if grads is not None else None
        ops.batchnorm_bwd(dz1, act1, y, stats, bn.weight.detach(), bn.bias.detach(), dy, dgamma=dg, dbeta=db, dz2=dz2,
                          act2=act2, mask=mask)
        if grads is not None and self.grad_ready is not None:
            self.grad_ready(name + ".weight")
            self.grad_ready(name + ".bias")
        return dy


class Pix2PixGeneratorNet(_NetBase, _BatchNormMixin):
    """Pix2PixGenerator (model_architectures.py:9-63): 8-level U-Net of 4x4 stride-2 (transposed) convolutions,
    BatchNorm2d with batch statistics, dropout in three blocks. The reference's in-place activations make a block
    return cat(lrelu(x), up(...)) and hand relu(cat(...)) to the up-convolution (see oracle.pix2pix_generator_forward);
    here every encoder activation e_k is written twice -- lrelu(e_k) for the next down-convolution and relu(e_k)
    straight into the first half of level k's concatenation buffer -- and relu(d_{k+1}) into its second half."""

    # (outer_nc, inner_nc, kind) from the outside in
    BLOCKS = [(3, 64, "outermost"), (64, 128, "middle"), (128, 256, "middle"), (256, 512, "middle"),
              (512, 512, "dropout"), (512, 512, "dropout"), (512, 512, "dropout"), (512, 512, "innermost")]

    def __init__(self, module):
        super().__init__(module)
        self.levels = []
        blk, prefix = module.model, "model.model."
        for level, (outer, inner, kind) in enumerate(self.BLOCKS):
            seq = blk.model
            if kind == "outermost":
                idx = dict(down=0, sub=1, up=3)
            elif kind == "innermost":
                idx = dict(down=1, up=3, upnorm=4)
            else:
                idx = dict(down=1, downnorm=2, sub=3, up=5, upnorm=6)
            lv = {"kind": kind, "outer": outer, "inner": inner,
                  "down": prefix + str(idx["down"]), "up": prefix + str(idx["up"]),
                  "downnorm": prefix + str(idx["downnorm"]) if "downnorm" in idx else None,
                  "upnorm": prefix + str(idx["upnorm"]) if "upnorm" in idx else None,
                  "downnorm_m": seq[idx["downnorm"]] if "downnorm" in idx else None,
                  "upnorm_m": seq[idx["upnorm"]] if "upnorm" in idx else None}
            self._add(lv["down"], seq[idx["down"]], 4, 2, 1)
            self._add(lv["up"], seq[idx["up"]], 4, 2, 1, transposed=True, use_bias=(kind == "outermost"))
            self.levels.append(lv)
            if "sub" in idx:
                prefix += str(idx["sub"]) + ".model."
                blk = seq[idx["sub"]]
        import torch.distributed as dist
        self.dropout_rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.dropout_masks = None  # test hook: list of three uint8 masks [pixels][c] in execution order
        self.dropout_seeds_dev = None  # int64[3] on the device: per-step seeds of the three masks (Pix2PixTrainer)

    def forward(self, x):
        self.repack()
        B, C, H, W = x.shape
        if H % 256 or W % 256:
            raise RuntimeError("Pix2PixGenerator: the 8-level U-Net needs inputs that are multiples of 256 pixels")
        t = {"x": x}
        xin = ActBuf(B, H, W, 16, zero=False)
        ops.pack_nchw(x, xin, 0, zero_rest=True)
        t["xin"] = xin
        cats = []  # cats[k]: input of level k's up-convolution = [relu(e_k) | relu(d_{k+1})]
        cur = xin
        for k, lv in enumerate(self.levels):
            h, w, c = cur.h // 2, cur.w // 2, lv["inner"]
            if lv["kind"] == "innermost":
                r7 = self._conv(cur, lv["down"], act=ACT_RELU)  # only consumer: relu(e_7) -> up-convolution
                t["r7"] = r7
                cats.append(r7)
                break
            cat = ActBuf(B, h, w, 2 * c, zero=False)
            cats.append(cat)
            if lv["kind"] == "outermost":
                a = self._conv(cur, lv["down"], act=ACT_LEAKY)  # a_0 = lrelu(e_0); relu(e_0) = relu(a_0)
                ops.batchnorm_apply(a, None, None, None, ACT_RELU, cat.channels(0, c))
                t["y0"] = a
            else:
                y = self._conv(cur, lv["down"])
                a = ActBuf(B, h, w, c, zero=False)
                t[f"s{k}"] = self._bn_forward(y, lv["downnorm_m"], ACT_LEAKY, a, ACT_RELU, cat.channels(0, c))
                t[f"y{k}"] = y
            t[f"a{k}"] = a
            cur = a
        t["cats"] = cats
        masks = []
        for k in range(7, 0, -1):  # up path: level k writes relu(d_k) into the second half of cats[k - 1]
            lv = self.levels[k]
            u = self._conv(cats[k], lv["up"])
            mask = None
            if lv["kind"] == "dropout":
                if self.dropout_masks is not None:
                    mask = self.dropout_masks[len(masks)]
                elif self.dropout_seeds_dev is not None:
                    # fused trainer: the step's seeds live on the device (written by the host before every step /
                    # graph replay from the same draws as below)
                    mask = torch.empty(u.n * u.h * u.w * u.c, dtype=torch.uint8, device=x.device)
                    ops.dropout_mask_dev(mask, self.dropout_seeds_dev[len(masks):len(masks) + 1], self.dropout_rank)
                else:
                    mask = torch.empty(u.n * u.h * u.w * u.c, dtype=torch.uint8, device=x.device)
                    # one draw from torch's global generator per mask: torch.manual_seed(47) before a forward (what the
                    # reference does before every evaluation-time generator call, model.py:393,497,579) reproduces the
                    # masks whatever ran before; under data parallelism every rank is seeded alike, so the rank is
                    # mixed in to give each shard its own masks
                    seed = int(torch.empty((), dtype=torch.int64).random_().item())
                    ops.dropout_mask(mask, seed * 1000003 + self.dropout_rank)
                masks.append(mask)
            c = lv["outer"]
            t[f"us{k}"] = self._bn_forward(u, lv["upnorm_m"], ACT_RELU, cats[k - 1].channels(c, c), mask=mask)
            t[f"u{k}"] = u
            t[f"m{k}"] = mask
        t["o"] = self._conv(cats[0], self.levels[0]["up"], fp32=True, act=ACT_TANH)
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
        ops.unpack_nchw(t["o"], out, 0)
        return out, t

    def backward(self, t, grads, dout_nchw, need_dx=False):
        cats = t["cats"]
        o = t["o"]
        dpre = ActBuf(o.n, o.h, o.w, 16, zero=False)
        ops.tanh_bwd_pack(dout_nchw, o, dpre)
        dcat = self._conv_bwd(self.levels[0]["up"], cats[0], dpre, grads, True)  # [.., 2 * 64]
        for k in range(1, 8):  # down the up path: gradient w.r.t. relu(d_k) is the second half of dcat_{k-1}
            lv = self.levels[k]
            c = lv["outer"]
            du = self._bn_backward(lv["upnorm"], lv["upnorm_m"], dcat.channels(c, c), ACT_RELU, t[f"u{k}"],
                                   t[f"us{k}"], grads, mask=t[f"m{k}"])
            dcat_next = self._conv_bwd(lv["up"], cats[k], du, grads, True)
            t[f"dcat{k - 1}"] = dcat
            dcat = dcat_next
        # innermost: dcat is the gradient w.r.t. r7 = relu(e_7)
        lv = self.levels[7]
        de = ActBuf(dcat.n, dcat.h, dcat.w, dcat.c, zero=False)
        ops.act_bwd(dcat, t["r7"], ACT_RELU, de)
        da = self._conv_bwd(lv["down"], t["a6"], de, grads, True)
        for k in range(6, 0, -1):  # e_k has two consumers: lrelu -> down conv (da), relu -> skip (first half of dcat_k)
            lv = self.levels[k]
            c = lv["inner"]
            dy = self._bn_backward(lv["downnorm"], lv["downnorm_m"], da, ACT_LEAKY, t[f"y{k}"], t[f"s{k}"], grads,
                                   dz2=t[f"dcat{k}"].channels(0, c), act2=ACT_RELU)
            da = self._conv_bwd(lv["down"], t[f"a{k - 1}"], dy, grads, True)
        lv = self.levels[0]
        de0 = ActBuf(da.n, da.h, da.w, da.c, zero=False)
        ops.batchnorm_bwd(da, ACT_LEAKY, t["y0"], None, None, None, de0, dz2=t["dcat0"].channels(0, lv["inner"]),
                          act2=ACT_RELU)
        dxin = self._conv_bwd(lv["down"], t["xin"], de0, grads, need_dx)
        if not need_dx:
            return None
        c_in = self.layers[lv["down"]].c_valid
        dx = torch.zeros(dxin.n, c_in, dxin.h, dxin.w, dtype=torch.float32, device=dxin.t.device)
        ops.unpack_nchw(dxin, dx, 0)
        return dx


class PatchGANBatchNormNet(_NetBase, _BatchNormMixin):
    """Pix2PixDiscriminator (model_architectures.py:65-85): 70x70 PatchGAN with BatchNorm2d (batch statistics) and
    no bias on the normalised convolutions."""

    def __init__(self, module):
        super().__init__(module)
        seq = module.model
        self._add("model.0", seq[0], 4, 2, 1, use_bias=True)
        self._add("model.2", seq[2], 4, 2, 1)
        self._add("model.5", seq[5], 4, 2, 1)
        self._add("model.8", seq[8], 4, 1, 1)
        self._add("model.11", seq[11], 4, 1, 1, use_bias=True)
        self.norms = {"model.3": seq[3], "model.6": seq[6], "model.9": seq[9]}
        self._add_s2d_stem("model.0")

    def forward_buf(self, din, xs=None):
        """din: ActBuf [B, H, W, 16] (channels beyond the real ones zero); xs: its space-to-depth copy if the caller
        kept it (tape["din_s2d"]). Returns (logits ActBuf fp32, tape)."""
        self.repack()
        t = {"din": din}
        t["a1"], t["din_s2d"] = self._stem_conv(din, "model.0", ACT_LEAKY, xs)
        cur = t["a1"]
        for i, (conv, norm) in enumerate((("model.2", "model.3"), ("model.5", "model.6"), ("model.8", "model.9")), 2):
            y = self._conv(cur, conv)
            a = ActBuf(y.n, y.h, y.w, y.c, zero=False)
            t[f"s{i}"] = self._bn_forward(y, self.norms[norm], ACT_LEAKY, a)
            t[f"y{i}"], t[f"a{i}"] = y, a
            cur = a
        t["logits"] = self._conv(cur, "model.11", fp32=True)
        return t["logits"], t

    def forward(self, x):
        B, C, H, W = x.shape
        din = ActBuf(B, H, W, 16, zero=False)
        ops.pack_nchw(x, din, 0, zero_rest=True)
        logits, t = self.forward_buf(din)
        return logits.to_nchw(1), t

    def backward(self, t, dlogits, grads, need_dx):
        d = self._conv_bwd("model.11", t["a4"], dlogits, grads, True)
        for i, (conv, norm) in zip((4, 3, 2), (("model.8", "model.9"), ("model.5", "model.6"), ("model.2", "model.3"))):
            dy = self._bn_backward(norm, self.norms[norm], d, ACT_LEAKY, t[f"y{i}"], t[f"s{i}"], grads)
            d = self._conv_bwd(conv, t[f"a{i - 1}"], dy, grads, True)
        dy1 = ActBuf(t["a1"].n, t["a1"].h, t["a1"].w, t["a1"].c, zero=False)
        ops.act_bwd(d, t["a1"], ACT_LEAKY, dy1)
        return self._conv_bwd("model.0", t["din"], dy1, grads, need_dx)


# ------------------------------------------------------------------------------------------------ segmentation U-Net
class UNetNet(_NetBase, _BatchNormMixin):
    """UNet(bilinear=False) forward (model_architectures.py:508-586) as used by calculate_metrics (model.py:380-400):
    3x3 zero-padded convolutions without bias + BatchNorm2d in TRAINING mode (the reference never calls .eval()) + ReLU,
    2x2 max pooling, 2x2 stride-2 transposed convolutions, channel concatenation, 1x1 head. Inference only."""

    def __init__(self, module):
        super().__init__(module)
        self.doubles = {}

        def double(prefix, seq):
            self._add(prefix + ".0", seq[0], 3, 1, 1)
            self._add(prefix + ".3", seq[3], 3, 1, 1)
            self.doubles[prefix] = (seq[1], seq[4])

        double("inc.double_conv", module.inc.double_conv)
        for i in range(1, 5):
            double(f"down{i}.maxpool_conv.1.double_conv", getattr(module, f"down{i}").maxpool_conv[1].double_conv)
        for i in range(1, 5):
            up = getattr(module, f"up{i}")
            self._add(f"up{i}.up", up.up, 2, 2, 0, transposed=True, use_bias=True)
            double(f"up{i}.conv.double_conv", up.conv.double_conv)
        self._add("outc.conv", module.outc.conv, 1, 1, 0, use_bias=True)

    def repack(self, force=False):
        # inference only: the dgrad layouts of the plain convolutions are never used, but the batched packer fills both
        super().repack(force)

    def _double(self, prefix, x, out=None):
        """conv-BN-ReLU twice; the second result may be written into `out` (a channel slice of a concatenation
        buffer)"""
        bn1, bn2 = self.doubles[prefix]
        y, st = self._conv_stats(x, prefix + ".0", batch=True, eps=bn1.eps)
        a = ActBuf(y.n, y.h, y.w, y.c, zero=False)
        self._bn_forward(y, bn1, ACT_RELU, a, stats=st)
        y2, st2 = self._conv_stats(a, prefix + ".3", batch=True, eps=bn2.eps)
        z = out if out is not None else ActBuf(y2.n, y2.h, y2.w, y2.c, zero=False)
        self._bn_forward(y2, bn2, ACT_RELU, z, stats=st2)
        return z

    def forward(self, x):
        """x: fp32 NCHW [B, 3, H, W] with H, W multiples of 16 -> (logits fp32 NCHW [B, 1, H, W], tape)"""
        self.repack()
        B, C, H, W = x.shape
        if H % 16 or W % 16:
            raise RuntimeError("UNet: input extent must be a multiple of 16")
        xin = ActBuf(B, H, W, 16, zero=False)
        ops.pack_nchw(x, xin, 0, zero_rest=True)
        widths = (64, 128, 256, 512, 1024)
        # cats[i]: input of up-level i's DoubleConv = [skip x_{i} | upsampled]; the encoder writes its output straight
        # into the first half
        cats = [ActBuf(B, H >> i, W >> i, 2 * widths[i], zero=False) for i in range(4)]
        cur = self._double("inc.double_conv", xin, out=cats[0].channels(0, 64))
        for i in range(1, 5):
            pooled = ActBuf(B, cur.h // 2, cur.w // 2, cur.c, zero=False)
            ops.maxpool2(cur, pooled)
            out = cats[i].channels(0, widths[i]) if i < 4 else None
            cur = self._double(f"down{i}.maxpool_conv.1.double_conv", pooled, out=out)
        for i in range(1, 5):
            lvl = 4 - i
            L = self.layers[f"up{i}.up"]
            ops.conv_dgrad(cur, L.spec, cats[lvl].channels(widths[lvl], widths[lvl]), bias=L.bias_pad)
            cur = self._double(f"up{i}.conv.double_conv", cats[lvl])
        logits = self._conv(cur, "outc.conv", fp32=True)
        return logits.to_nchw(1), {"logits": logits}
