"""Fused training steps over the native executors (reference: Model.train_paired / train_cycle inner loops,
models/model.py:611-651 and :678-751).

The step schedule, loss definitions, optimiser and update order are the reference's; what changes is the
execution: no autograd graph, whole-network kernels, parameters / gradients / Adam moments in flat fp32 buffers
(one Adam launch per network, one NCCL all-reduce per gradient bucket), losses kept on the device.
Data parallelism: one process per GPU, each rank steps on its shard of the global batch; gradients are summed
with NCCL over NVLink (bucketed, overlapped with the generator backward) and scaled by 1/world_size inside Adam.
"""
import os

import torch
import torch.distributed as dist

from . import networks, ops
from .ops import ActBuf


class FlatParams:
    """Re-homes a module's parameters into one flat fp32 buffer (each Parameter becomes a view), plus matching
    flat gradient and Adam-moment buffers. state_dict() / load_state_dict() keep working on the views."""

    def __init__(self, module):
        self.module = module
        self.named = list(module.named_parameters())
        total = sum(p.numel() for _, p in self.named)
        dev = self.named[0][1].device
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = {}
        for n, p in self.named:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            self.offsets[n] = (off, k)
            off += k
        self.grads = networks.Grads(self.named)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        # {step, lr (float bits), step_size, sqrt(bias correction 2)}: the step count and learning rate live on the
        # device so that the launch arguments of a training step never change (CUDA-graph replay)
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self._steps = 0
        self._lr = None

    @property
    def steps(self):
        """optimiser steps taken (host mirror of the device counter; checkpoints store it)"""
        return self._steps

    @steps.setter
    def steps(self, value):
        self._steps = int(value)
        self.state[0:1].fill_(self._steps)

    def set_lr(self, lr):
        if lr != self._lr:
            self.state[1:2].view(torch.float32).fill_(float(lr))
            self._lr = lr

    def adam(self, grad_scale=1.0, betas=(0.5, 0.999), eps=1e-8):
        """one Adam update; the device-side counter advances, the caller mirrors it with `note_step()`"""
        ops.adam_step_dev(self.flat, self.grads.flat, self.m, self.v, self.state, betas[0], betas[1], eps, grad_scale)

    def note_step(self):
        self._steps += 1


class _BucketReducer:
    """Sums a flat gradient buffer across ranks in contiguous buckets. Buckets are launched (async, on NCCL's own
    stream) as soon as every layer whose gradient lives in them has been produced, so the transfers overlap the
    rest of the backward pass."""

    def __init__(self, flat_params, bucket_bytes=None, group=None):
        if bucket_bytes is None:
            # FPG_DDP_BUCKET_MB: 0 (default) = one all-reduce after the backward pass. Measured inside the replayed
            # graph: 8 x B200 12140 tiles/s against 11934 with 8 MB buckets overlapped with the backward (2 x B200
            # 3116 / 3089): an NCCL kernel that overlaps a persistent one-CTA-per-SM conv kernel takes SMs from it and
            # the displaced CTAs run after the others (up to one extra tile time), which costs more than the 47 MB
            # all-reduce over NVLink does on its own.
            mb = float(os.environ.get("FPG_DDP_BUCKET_MB", "0"))
            bucket_bytes = int(mb * (1 << 20)) if mb > 0 else 1 << 62
        self.fp = flat_params
        self.group = group
        self.bounds = []
        self.layer_bucket = {}
        start, cur = 0, 0
        # parameters are laid out in forward order; backward produces them roughly back to front
        for n, p in flat_params.named:
            off, k = flat_params.offsets[n]
            if (off + k - start) * 4 > bucket_bytes and off > start:
                self.bounds.append((start, off))
                start = off
                cur += 1
            self.layer_bucket[n] = cur
        self.bounds.append((start, flat_params.flat.numel()))
        self.pending = None
        self.works = []

    def start(self):
        self.pending = [0] * len(self.bounds)
        for n in self.layer_bucket:
            self.pending[self.layer_bucket[n]] += 1
        self.works = []

    def ready(self, name):
        b = self.layer_bucket[name]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            s, e = self.bounds[b]
            self.works.append(dist.all_reduce(self.fp.grads.flat[s:e], group=self.group, async_op=True))

    def finish(self):
        for b, left in enumerate(self.pending):
            if left > 0:  # parameters whose gradient is never produced (biases before an InstanceNorm)
                s, e = self.bounds[b]
                self.works.append(dist.all_reduce(self.fp.grads.flat[s:e], group=self.group, async_op=True))
                self.pending[b] = 0
        for w in self.works:
            w.wait()
        self.works = []


class PairedTrainer:
    """One `train_paired` iteration (model.py:611-651) for the InstanceNorm paired model (PairedAttention)."""

    LOSS_KEYS = ("losses_discriminator_real", "losses_discriminator_synthetic", "losses_generator_synthetic",
                 "l1_losses_generator_synthetic")

    def __init__(self, generator, discriminator, world_size=1, group=None, l1_weight=100.0):
        self.gen_module, self.dis_module = generator, discriminator
        self.gp, self.dp = FlatParams(generator), FlatParams(discriminator)
        self.G, self.D = generator._executor(), discriminator._executor()
        self.world_size = world_size
        self.group = group
        self.l1_weight = l1_weight
        dev = self.gp.flat.device
        self.loss_buf = torch.zeros(4, dtype=torch.float32, device=dev)
        self.g_reducer = _BucketReducer(self.gp, group=group) if world_size > 1 else None
        self.launches = 0
        # CUDA-graph replay of the whole step (341 launches, the NCCL gradient reductions included): removes host
        # launch overhead and inter-kernel gaps (measured 5-6 % of the step). Captured NCCL work must be released before
        # the process group is destroyed (destroy_process_group() hangs otherwise): release_graphs(), also at exit.
        self.use_graph = os.environ.get("FPG_CUDA_GRAPH", "1") != "0" and (
            world_size == 1 or os.environ.get("FPG_CUDA_GRAPH_DDP", "1") == "1")
        self._graphs, self._eager_calls = {}, {}
        if world_size > 1:
            import atexit
            import weakref
            ref = weakref.ref(self)
            atexit.register(lambda: ref() is not None and ref().release_graphs())

    def _force_repack(self, net):
        net.repack(force=True)

    def step(self, input_stack, output_image, lr_g=0.0002, lr_d=0.0002):
        """input_stack [B,C,H,W], output_image [B,3,H,W]: fp32 CUDA tensors. Returns the generated image (fp32 NCHW;
        with graph replay it is a static buffer that the next step overwrites).
        The four losses of the step are left in self.loss_buf (device) in LOSS_KEYS order."""
        self.gp.set_lr(lr_g)
        self.dp.set_lr(lr_d)
        out = None
        if self.use_graph and ops.PROFILE is None:
            key = (tuple(input_stack.shape), tuple(output_image.shape))
            entry = self._graphs.get(key)
            if entry is None and self._eager_calls.get(key, 0) >= 2:
                entry = self._graphs[key] = self._capture(input_stack, output_image)
            if entry is not None:
                sx, sy, graph, out, launches = entry
                sx.copy_(input_stack, non_blocking=True)
                sy.copy_(output_image, non_blocking=True)
                graph.replay()
                ops.LAUNCHES += launches
            else:  # first calls run eagerly: workspaces, job tables and kernel attributes are set up lazily
                self._eager_calls[key] = self._eager_calls.get(key, 0) + 1
        if out is None:
            out = self._step_impl(input_stack, output_image)
        self.gp.note_step()
        self.dp.note_step()
        return out

    def release_graphs(self):
        """drop the captured steps (and their memory pool); the next step() calls run eagerly and re-capture"""
        self._graphs.clear()
        self._eager_calls.clear()

    def _capture(self, input_stack, output_image):
        sx, sy = input_stack.clone(), output_image.clone()
        before = ops.LAUNCHES
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._step_impl(sx, sy)
        launches = ops.LAUNCHES - before
        ops.LAUNCHES = before  # capturing records the launches, it does not run them
        return sx, sy, graph, out, launches

    def _step_impl(self, input_stack, output_image):
        G, D = self.G, self.D
        B, C, H, W = input_stack.shape
        inv_w = 1.0 / self.world_size
        # D input for [synthetic | real] halves: channels 0..C-1 = input stack, C..C+2 = image (model.py:616-617)
        din = ActBuf(2 * B, H, W, 16, zero=False)
        fake, real = din.batch_slice(0, B), din.batch_slice(B, B)
        ops.pack_nchw(input_stack, fake, 0, zero_rest=True)
        ops.pack_nchw(input_stack, real, 0, zero_rest=True)
        ops.pack_nchw(output_image, real, C)
        synthetic, gtape = G.forward(input_stack, d_input=fake, d_c0=C)                         # :615

        # ---- discriminator update (:620-633): both halves in one pass (InstanceNorm is per sample)
        logits, dtape = D.forward_buf(din)
        dlog = ActBuf(2 * B, logits.h, logits.w, 16, zero=False)
        ops.mse_const_loss(logits.batch_slice(0, B), 0.0, 1.0, 0.5, self.loss_buf[1:2], dlog.batch_slice(0, B))
        ops.mse_const_loss(logits.batch_slice(B, B), 1.0, 1.0, 0.5, self.loss_buf[0:1], dlog.batch_slice(B, B))
        D.backward(dtape, dlog, self.dp.grads, need_dx=False)
        if self.world_size > 1:
            dist.all_reduce(self.dp.grads.flat, group=self.group)
        self.dp.adam(grad_scale=inv_w)
        self._force_repack(D)

        # ---- generator update (:636-646): adversarial term through the UPDATED discriminator
        logits_g, dtape_g = D.forward_buf(fake)
        dlog_g = ActBuf(B, logits_g.h, logits_g.w, 16, zero=False)
        ops.mse_const_loss(logits_g, 1.0, 1.0, 1.0, self.loss_buf[2:3], dlog_g)
        d_din = D.backward(dtape_g, dlog_g, None, need_dx=True)
        dl1 = torch.empty_like(synthetic)
        ops.l1_loss(synthetic, output_image, self.l1_weight, 1.0, self.loss_buf[3:4], dpred=dl1)
        if self.g_reducer is not None:
            self.g_reducer.start()
            G.grad_ready = self.g_reducer.ready
        G.backward(gtape, self.gp.grads, dout_nchw=dl1, dout_nhwc=d_din, dout_c0=C, need_dx=False)
        if self.g_reducer is not None:
            G.grad_ready = None
            self.g_reducer.finish()
        self.gp.adam(grad_scale=inv_w)
        self._force_repack(G)
        return synthetic

    def losses(self):
        """Host copy of the last step's losses (one device->host sync), averaged over ranks."""
        buf = self.loss_buf.clone()
        if self.world_size > 1:
            dist.all_reduce(buf, group=self.group)
            buf /= self.world_size
        vals = buf.tolist()
        return dict(zip(self.LOSS_KEYS, vals))
