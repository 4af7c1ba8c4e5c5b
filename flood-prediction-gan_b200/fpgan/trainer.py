"""Fused training steps over the native executors (reference: Model.train_paired / train_cycle inner loops,
models/model.py:611-651 and :678-751).

The step schedule, loss definitions, optimiser and update order are the reference's; what changes is the
execution: no autograd graph, whole-network kernels, parameters / gradients / Adam moments in flat fp32 buffers
(one Adam launch per network, one NCCL all-reduce per gradient bucket), losses kept on the device.
Data parallelism: one process per GPU, each rank steps on its shard of the global batch; gradients are summed
with NCCL over NVLink (bucketed, overlapped with the generator backward) and scaled by 1/world_size inside Adam.
"""
import os
import random

import torch
import torch.distributed as dist

from . import networks, ops, peer
from .ops import ActBuf


class FlatParams:
    """Re-homes the parameters of one module -- or of several modules that share an optimiser (train_cycle chains both
    generators / both discriminators into one Adam, model.py:112-122) -- into one flat fp32 buffer (each Parameter
    becomes a view), plus matching flat gradient and Adam-moment buffers: one Adam launch and one all-reduce per
    optimiser. state_dict() / load_state_dict() keep working on the views. With several modules the names in
    `named` / `offsets` are prefixed "<module index>."; `grads_of[i]` addresses module i's gradients by its own names."""

    def __init__(self, *modules):
        self.modules = modules
        self.module = modules[0]
        if len(modules) == 1:
            self.named = list(modules[0].named_parameters())
        else:
            self.named = [(f"{i}.{n}", p) for i, m in enumerate(modules) for n, p in m.named_parameters()]
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
        self.grads_of = self._per_module(self.grads.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        # {step, lr (float bits), step_size, sqrt(bias correction 2)}: the step count and learning rate live on the
        # device so that the launch arguments of a training step never change (CUDA-graph replay)
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self._steps = 0
        self._lr = None

    def _per_module(self, flat):
        out, off = [], 0
        for m in self.modules:
            named = list(m.named_parameters())
            size = sum(p.numel() for _, p in named)
            out.append(networks.Grads(named, flat=flat[off:off + size]))
            off += size
        return out

    def extra_grads(self):
        """a second flat gradient buffer (+ per-module views) for a network that runs more than once per step; the
        caller sums it into `grads.flat` (ops.add_f32) before the all-reduce / Adam"""
        flat = torch.zeros_like(self.grads.flat)
        return flat, self._per_module(flat)

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

    def adam(self, grad_scale, betas=(0.5, 0.999), eps=1e-8, fused=None):
        if fused is not None:
            fused.run(grad_scale, betas=betas, eps=eps)
        else:
            self.fp.adam(grad_scale=grad_scale, betas=betas, eps=eps)

    def close(self):
        pass


def optimiser_step(fp, reducer, fused, nets, grad_scale):
    """Adam on `fp` (through the data-parallel reducer when there is one) + the bf16 operands of `nets` rebuilt: one
    launch with `fused` (networks.AdamPack), else the update followed by the batched repack."""
    if reducer is not None:
        reducer.adam(grad_scale, fused=fused)
    elif fused is not None:
        fused.run(grad_scale)
    else:
        fp.adam(grad_scale=grad_scale)
    if fused is None:
        for net in nets:
            net.repack(force=True)


def make_reducers(world_size, group, *flat_params, overlappable=True):
    """One gradient reducer per optimiser for data-parallel training, None each for a single process.
    FPG_DDP=peer: the copy-engine exchange over peer memory (peer.PeerReducer; needs every rank to map the others'
    memory: one NVSwitch / NVLink box); FPG_DDP=nccl: NCCL all-reduces (_BucketReducer). Default (auto): the peer
    exchange where the backward pass can hide its transfers (`overlappable`: the paired steps) AND the world is small
    (<= 4 ranks) -- its cost grows with the world size (W - 1 pushes per bucket, W gradient sources read by Adam:
    measured on 8 x B200 10.01 ms per step against 9.96 ms with NCCL's in-switch reduction, on 2 x B200 9.66-9.77 ms
    against 9.79-9.86 ms), while an NVLS all-reduce costs about the same at any size. Where nothing can be hidden
    (train_cycle: every network runs 2-3 times per step and its gradient is complete only after the last run) an
    all-reduce is the right collective -- it moves 2(W-1)/W of the buffer per rank, the all-gather W-1 times it."""
    if world_size <= 1:
        return [None] * len(flat_params)
    mode = os.environ.get("FPG_DDP", "auto")
    want_peer = mode == "peer" or (mode == "auto" and overlappable and world_size <= 4)
    if want_peer and peer.supported(group):
        return [peer.PeerReducer(fp, group=group) for fp in flat_params]
    return [_BucketReducer(fp, group=group) for fp in flat_params]


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
        self.g_reducer, self.d_reducer = make_reducers(world_size, group, self.gp, self.dp)
        # FPG_ADAM_PACK=0: Adam and the operand repack as two launches (cross-check of the fused kernel)
        fuse = os.environ.get("FPG_ADAM_PACK", "1") != "0"
        self.g_fused = networks.AdamPack(self.gp, [self.G]) if fuse else None
        self.d_fused = networks.AdamPack(self.dp, [self.D]) if fuse else None
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

    def close(self):
        """collective teardown of the data-parallel state: captured steps, then the peer-memory mappings"""
        self.release_graphs()
        for r in (self.g_reducer, self.d_reducer):
            if r is not None:
                r.close()
        self.g_reducer = self.d_reducer = None

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

    def _backward_reduced(self, net, reducer, run):
        """run() = a backward pass of executor `net`; with a reducer its gradient buckets are sent to the other ranks
        as they complete and the call returns once the exchange is in flight / has landed (reducer.finish)"""
        if reducer is None:
            return run()
        reducer.start()
        net.grad_ready = reducer.ready
        try:
            out = run()
        finally:
            net.grad_ready = None
        reducer.finish()
        return out

    def _phase_d(self, input_stack, output_image, reducer=None):
        """generator forward + discriminator forward / backward on [synthetic | real] (:615-632): leaves the
        discriminator gradients of this batch in self.dp.grads and returns what the generator phase needs"""
        G, D = self.G, self.D
        B, C, H, W = input_stack.shape
        # D input for [synthetic | real] halves: channels 0..C-1 = input stack, C..C+2 = image (model.py:616-617)
        din = ActBuf(2 * B, H, W, 16, zero=False)
        fake, real = din.batch_slice(0, B), din.batch_slice(B, B)
        if input_stack.is_contiguous() and output_image.is_contiguous():
            gin = ActBuf(B, H, W, 16, halo=3, zero=False)
            ops.pack_paired_inputs(input_stack, output_image, gin, fake, real)  # one pass over the fp32 batch
        else:
            gin = None
            ops.pack_nchw(input_stack, fake, 0, zero_rest=True)
            ops.pack_nchw(input_stack, real, 0, zero_rest=True)
            ops.pack_nchw(output_image, real, C)
        synthetic, gtape = G.forward(input_stack, d_input=fake, d_c0=C, xin=gin)                # :615

        # ---- discriminator update (:620-633): both halves in one pass (InstanceNorm is per sample)
        logits, dtape = D.forward_buf(din)
        dlog = ActBuf(2 * B, logits.h, logits.w, 16, zero=False)
        ops.mse_const_loss(logits.batch_slice(0, B), 0.0, 1.0, 0.5, self.loss_buf[1:2], dlog.batch_slice(0, B))
        ops.mse_const_loss(logits.batch_slice(B, B), 1.0, 1.0, 0.5, self.loss_buf[0:1], dlog.batch_slice(B, B))
        self._backward_reduced(D, reducer, lambda: D.backward(dtape, dlog, self.dp.grads, need_dx=False))
        xs = dtape.get("din_s2d")  # space-to-depth copy of [synthetic | real]: its first half serves the generator phase
        return synthetic, gtape, fake, output_image, C, xs.batch_slice(0, B) if xs is not None else None

    def _phase_g(self, state, reducer=None):
        """generator update terms (:636-645) through the UPDATED discriminator: leaves the generator gradients of
        this batch in self.gp.grads"""
        G, D = self.G, self.D
        synthetic, gtape, fake, output_image, C, xs_fake = state
        B = synthetic.shape[0]
        logits_g, dtape_g = D.forward_buf(fake, xs_fake)
        dlog_g = ActBuf(B, logits_g.h, logits_g.w, 16, zero=False)
        ops.mse_const_loss(logits_g, 1.0, 1.0, 1.0, self.loss_buf[2:3], dlog_g)
        d_din = D.backward(dtape_g, dlog_g, None, need_dx=True)
        dl1 = torch.empty_like(synthetic)
        ops.l1_loss(synthetic, output_image, self.l1_weight, 1.0, self.loss_buf[3:4], dpred=dl1)
        self._backward_reduced(G, reducer, lambda: G.backward(gtape, self.gp.grads, dout_nchw=dl1, dout_nhwc=d_din,
                                                              dout_c0=C, need_dx=False))

    def _step_impl(self, input_stack, output_image):
        inv_w = 1.0 / self.world_size
        state = self._phase_d(input_stack, output_image, self.d_reducer)
        optimiser_step(self.dp, self.d_reducer, self.d_fused, [self.D], inv_w)
        self._phase_g(state, self.g_reducer)
        optimiser_step(self.gp, self.g_reducer, self.g_fused, [self.G], inv_w)
        return state[0]

    def step_accumulated(self, shards, lr_g=0.0002, lr_d=0.0002):
        """One optimiser step over several micro-batches [(input_stack, output_image), ...] with the gradients summed
        in shard order and averaged -- what W data-parallel ranks compute together, in ONE process and with the very
        same kernels per shard (plans depend on the per-process batch). The data-parallel parity check compares the
        NCCL run against this; it is also plain gradient accumulation for batches that do not fit. Eager (no graph).
        Returns the per-shard generated images; loss_buf holds the shard-averaged losses."""
        assert self.world_size == 1, "step_accumulated emulates the ranks in a single process"
        self.gp.set_lr(lr_g)
        self.dp.set_lr(lr_d)
        n = len(shards)
        states, loss_sum = [], torch.zeros_like(self.loss_buf)
        dsum = torch.zeros_like(self.dp.grads.flat)
        for x, y in shards:
            states.append(self._phase_d(x, y))
            ops.add_f32(dsum, self.dp.grads.flat)
            loss_sum[0:2] += self.loss_buf[0:2]
        self.dp.grads.flat.copy_(dsum)
        optimiser_step(self.dp, None, self.d_fused, [self.D], 1.0 / n)
        gsum = torch.zeros_like(self.gp.grads.flat)
        for st in states:
            self._phase_g(st)
            ops.add_f32(gsum, self.gp.grads.flat)
            loss_sum[2:4] += self.loss_buf[2:4]
        self.gp.grads.flat.copy_(gsum)
        optimiser_step(self.gp, None, self.g_fused, [self.G], 1.0 / n)
        self.loss_buf.copy_(loss_sum / n)
        self.gp.note_step()
        self.dp.note_step()
        return [st[0] for st in states]

    def losses(self):
        """Host copy of the last step's losses (one device->host sync), averaged over ranks."""
        buf = self.loss_buf.clone()
        if self.world_size > 1:
            dist.all_reduce(buf, group=self.group)
            buf /= self.world_size
        vals = buf.tolist()
        return dict(zip(self.LOSS_KEYS, vals))


class Pix2PixTrainer(PairedTrainer):
    """One `train_paired` iteration (model.py:611-651) for Pix2Pix (BatchNorm U-Net generator with dropout + BatchNorm
    PatchGAN, model_architectures.py:9-85), fused like PairedTrainer: no autograd graph, native loss kernels, flat-buffer
    Adam, CUDA-graph replay, gradient exchange per optimiser under data parallelism. Differences from the InstanceNorm
    pair, all the reference's: BatchNorm uses BATCH statistics, so the discriminator's synthetic and real passes are two
    separate calls (two running-statistics updates, in the reference's order: synthetic first) whose parameter gradients
    are summed, and a third call in the generator phase; the dropout masks of the generator change every step (seeds
    drawn on the host from torch's global generator -- one per mask, exactly as the module path draws them -- and
    written to the device before the step / replay); BatchNorm statistics and dropout masks are per rank (what
    torch DDP does without SyncBatchNorm; parity with the single-process reference is defined per replica)."""

    def __init__(self, generator, discriminator, world_size=1, group=None, l1_weight=100.0):
        super().__init__(generator, discriminator, world_size=world_size, group=group, l1_weight=l1_weight)
        self.d_extra_flat, d_extra = self.dp.extra_grads()
        self.d_extra = d_extra[0]
        dev = self.gp.flat.device
        self.seeds = torch.zeros(3, dtype=torch.int64, device=dev)
        self.inject_masks = None  # test hook: three uint8 keep-masks in execution order instead of drawn ones

    def _draw_seeds(self):
        vals = []
        for _ in range(3):  # the draws Pix2PixGeneratorNet.forward makes, in its order
            v = (int(torch.empty((), dtype=torch.int64).random_().item()) * 1000003) & (2 ** 64 - 1)
            vals.append(v - 2 ** 64 if v >= 2 ** 63 else v)
        self.seeds.copy_(torch.tensor(vals, dtype=torch.int64))

    def step(self, input_stack, output_image, lr_g=0.0002, lr_d=0.0002):
        self.G.dropout_masks = self.inject_masks
        self.G.dropout_seeds_dev = self.seeds
        if self.inject_masks is None:
            self._draw_seeds()
        try:
            return super().step(input_stack, output_image, lr_g=lr_g, lr_d=lr_d)
        finally:
            self.G.dropout_seeds_dev = None  # module-path calls (evaluation) keep drawing on the host

    def _phase_d(self, input_stack, output_image, reducer=None):
        G, D = self.G, self.D
        B, C, H, W = input_stack.shape
        synthetic, gtape = G.forward(input_stack)                                               # :615
        fake = ActBuf(B, H, W, 16, zero=False)
        real = ActBuf(B, H, W, 16, zero=False)
        ops.pack_nchw(input_stack, fake, 0, zero_rest=True)                                     # torch.cat, :616-617
        ops.pack_nchw(synthetic, fake, C)
        ops.pack_nchw(input_stack, real, 0, zero_rest=True)
        ops.pack_nchw(output_image, real, C)
        # ---- discriminator update (:620-633): synthetic pass, then real pass (BatchNorm statistics per call)
        lf, tf = D.forward_buf(fake)
        lr_, tr_ = D.forward_buf(real)
        dlf = ActBuf(B, lf.h, lf.w, 16, zero=False)
        dlr = ActBuf(B, lr_.h, lr_.w, 16, zero=False)
        ops.mse_const_loss(lf, 0.0, 1.0, 0.5, self.loss_buf[1:2], dlf)
        ops.mse_const_loss(lr_, 1.0, 1.0, 0.5, self.loss_buf[0:1], dlr)
        D.backward(tf, dlf, self.d_extra, need_dx=False)

        D.backward(tr_, dlr, self.dp.grads, need_dx=False)
        ops.add_f32(self.dp.grads.flat, self.d_extra_flat)
        if reducer is not None:  # complete only after the sum of the two passes: one exchange, nothing to overlap with
            reducer.start()
            reducer.finish()
        return synthetic, gtape, fake, output_image, C, tf.get("din_s2d")

    def _phase_g(self, state, reducer=None):
        G, D = self.G, self.D
        synthetic, gtape, fake, output_image, C, xs_fake = state
        B = synthetic.shape[0]
        lg, tg = D.forward_buf(fake, xs_fake)                                                   # :637, updated D
        dlg = ActBuf(B, lg.h, lg.w, 16, zero=False)
        ops.mse_const_loss(lg, 1.0, 1.0, 1.0, self.loss_buf[2:3], dlg)
        d_din = D.backward(tg, dlg, None, need_dx=True)
        dout = torch.empty_like(synthetic)
        ops.l1_loss(synthetic, output_image, self.l1_weight, 1.0, self.loss_buf[3:4], dpred=dout)
        ops.unpack_nchw(d_din, dout, c0=C, accumulate=True)  # + the adversarial gradient w.r.t. the synthetic image
        self._backward_reduced(G, reducer, lambda: G.backward(gtape, self.gp.grads, dout, need_dx=False))


class _History:
    """Host side of the history buffer of generated images (get_buffer_image, model.py:275-294): the reference's
    decisions, drawn from Python's global `random` in the reference's order, as {use_slot, store_slot} for the
    device-resident pool (ops.history_exchange). The first 50 images are stored and returned as they are; afterwards
    with probability 0.5 a random stored image is returned and replaced by the new one."""

    SIZE = 50

    def __init__(self):
        self.count = 0

    def decide(self):
        if self.count < self.SIZE:
            self.count += 1
            return -1, self.count - 1
        if random.uniform(0, 1) > 0.5:
            idx = random.randint(0, self.SIZE - 1)
            return idx, idx
        return -1, -1


class CycleTrainer:
    """One `train_cycle` iteration (model.py:678-739) for CycleGAN / AttentionGAN, fused like PairedTrainer: no autograd
    graph, native loss kernels, flat-buffer Adam over the chained generators / discriminators, device-resident history
    buffers, one gradient all-reduce per optimiser phase, the whole step replayed as a CUDA graph.

    Schedule of the reference: four generator passes (synthetic post / pre from the real images, recreated post / pre
    from the synthetic ones with the topography conditions re-attached, :680-689), optional identity passes (:701-702),
    generator update from the LSGAN terms through the frozen discriminators + 10 x cycle L1 (+ 5 x identity L1)
    (:703-713), then the discriminator update on the real images and on images drawn from the history buffers
    (:722-735). Each generator runs 2-3 times per step: every run writes its parameter gradients into its own flat
    buffer, summed before Adam."""

    def __init__(self, pre_to_post, post_to_pre, pre_discriminator, post_discriminator, add_identity_loss=False,
                 world_size=1, group=None):
        self.identity = add_identity_loss
        # optimiser parameter order of the reference: generators (pre_to_post, post_to_pre), discriminators (post, pre)
        self.gp = FlatParams(pre_to_post, post_to_pre)
        self.dp = FlatParams(post_discriminator, pre_discriminator)
        self.Gpp, self.Gpr = pre_to_post._executor(), post_to_pre._executor()
        self.Dpost, self.Dpre = post_discriminator._executor(), pre_discriminator._executor()
        self.g_extra = [self.gp.extra_grads() for _ in range(2 if add_identity_loss else 1)]
        self.world_size, self.group = world_size, group
        self.g_reducer, self.d_reducer = make_reducers(world_size, group, self.gp, self.dp, overlappable=False)
        fuse = os.environ.get("FPG_ADAM_PACK", "1") != "0"
        self.g_fused = networks.AdamPack(self.gp, [self.Gpp, self.Gpr]) if fuse else None
        self.d_fused = networks.AdamPack(self.dp, [self.Dpost, self.Dpre]) if fuse else None
        self.loss_keys = self.LOSS_KEYS + (self.IDENTITY_KEYS if add_identity_loss else ())
        dev = self.gp.flat.device
        self.loss_buf = torch.zeros(len(self.loss_keys), dtype=torch.float32, device=dev)
        self.hist_pre, self.hist_post = _History(), _History()
        self.hist_ctrl = torch.full((4,), -1, dtype=torch.int32, device=dev)  # {pre use, pre store, post use, post store}
        self._pools = {}
        self.use_graph = os.environ.get("FPG_CUDA_GRAPH", "1") != "0" and (
            world_size == 1 or os.environ.get("FPG_CUDA_GRAPH_DDP", "1") == "1")
        self._graphs, self._eager_calls = {}, {}
        if world_size > 1:
            import atexit
            import weakref
            ref = weakref.ref(self)
            atexit.register(lambda: ref() is not None and ref().release_graphs())

    LOSS_KEYS = ("losses_generator_post", "losses_generator_pre", "losses_pre_to_post_cycle",
                 "losses_post_to_pre_cycle", "losses_discriminator_pre_real", "losses_discriminator_post_real",
                 "losses_discriminator_pre_synthetic", "losses_discriminator_post_synthetic")
    IDENTITY_KEYS = ("losses_identity_post", "losses_identity_pre")

    def release_graphs(self):
        self._graphs.clear()
        self._eager_calls.clear()

    def close(self):
        """collective teardown of the data-parallel state: captured steps, then the peer-memory mappings"""
        self.release_graphs()
        for r in (self.g_reducer, self.d_reducer):
            if r is not None:
                r.close()
        self.g_reducer = self.d_reducer = None

    def _pool(self, which, like):
        """device pool of 50 packed discriminator inputs [50][B][H][W][16] bf16 for this batch shape"""
        key = (which, tuple(like.t.shape))
        if key not in self._pools:
            self._pools[key] = torch.empty((_History.SIZE,) + tuple(like.t.shape), dtype=torch.bfloat16,
                                           device=like.t.device)
        return self._pools[key]

    def step(self, input_stack, output_image, lr_g=0.0002, lr_d=0.0002):
        """input_stack [B,C,H,W] (pre-flood RGB + topography conditions), output_image [B,3,H,W]: fp32 CUDA tensors.
        Returns (synthetic_post, synthetic_pre) [B,3,H,W]; the losses are left in self.loss_buf in loss_keys order."""
        self.gp.set_lr(lr_g)
        self.dp.set_lr(lr_d)
        # the reference draws the pre buffer's decision first, then the post buffer's (model.py:719-720)
        ctrl = list(self.hist_pre.decide()) + list(self.hist_post.decide())
        self.hist_ctrl.copy_(torch.tensor(ctrl, dtype=torch.int32))
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
            else:
                self._eager_calls[key] = self._eager_calls.get(key, 0) + 1
        if out is None:
            out = self._step_impl(input_stack, output_image)
        self.gp.note_step()
        self.dp.note_step()
        return out

    def _capture(self, input_stack, output_image):
        sx, sy = input_stack.clone(), output_image.clone()
        self._pool("pre", ActBuf(sx.shape[0], sx.shape[2], sx.shape[3], 16, zero=False))  # allocate outside the capture
        self._pool("post", ActBuf(sx.shape[0], sx.shape[2], sx.shape[3], 16, zero=False))
        before = ops.LAUNCHES
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._step_impl(sx, sy)
        launches = ops.LAUNCHES - before
        ops.LAUNCHES = before
        return sx, sy, graph, out, launches

    def _step_impl(self, x_pre, y_post):
        inv_w = 1.0 / self.world_size
        state = self._phase_g(x_pre, y_post)
        # every network runs 2-3 times per step and its gradient is complete only after the last sum: one exchange
        # per optimiser phase (the step is 3.3 x longer per exchanged byte than the paired one)
        if self.g_reducer is not None:
            self.g_reducer.start()
            self.g_reducer.finish()
        optimiser_step(self.gp, self.g_reducer, self.g_fused, [self.Gpp, self.Gpr], inv_w)
        self._phase_d(state)
        if self.d_reducer is not None:
            self.d_reducer.start()
            self.d_reducer.finish()
        optimiser_step(self.dp, self.d_reducer, self.d_fused, [self.Dpre, self.Dpost], inv_w)
        return state[0], state[1]

    def step_accumulated(self, shards, lr_g=0.0002, lr_d=0.0002):
        """As PairedTrainer.step_accumulated: one optimiser step over several micro-batches, gradients summed in shard
        order -- what W data-parallel ranks compute together, in one process with the same per-shard kernels. Meant
        for the first 50 steps, while the history buffers return the current images (every rank has its own buffer)."""
        assert self.world_size == 1
        self.gp.set_lr(lr_g)
        self.dp.set_lr(lr_d)
        ctrl = list(self.hist_pre.decide()) + list(self.hist_post.decide())
        assert ctrl[0] < 0 and ctrl[2] < 0, "step_accumulated: the history buffers are past their fill phase"
        self.hist_ctrl.copy_(torch.tensor([-1, -1, -1, -1], dtype=torch.int32))  # nothing stored: one pool, many shards
        n = len(shards)
        states, loss_sum = [], torch.zeros_like(self.loss_buf)
        gsum = torch.zeros_like(self.gp.grads.flat)
        g_keys = [0, 1, 2, 3] + ([8, 9] if self.identity else [])
        for x, y in shards:
            states.append(self._phase_g(x, y))
            ops.add_f32(gsum, self.gp.grads.flat)
            loss_sum[g_keys] += self.loss_buf[g_keys]
        self.gp.grads.flat.copy_(gsum)
        optimiser_step(self.gp, None, self.g_fused, [self.Gpp, self.Gpr], 1.0 / n)
        dsum = torch.zeros_like(self.dp.grads.flat)
        for st in states:
            self._phase_d(st)
            ops.add_f32(dsum, self.dp.grads.flat)
            loss_sum[4:8] += self.loss_buf[4:8]
        self.dp.grads.flat.copy_(dsum)
        optimiser_step(self.dp, None, self.d_fused, [self.Dpre, self.Dpost], 1.0 / n)
        self.loss_buf.copy_(loss_sum / n)
        self.gp.note_step()
        self.dp.note_step()
        return [(st[0], st[1]) for st in states]

    def _phase_g(self, x_pre, y_post):
        """the generator passes and the generator update terms (:680-713): leaves the summed generator gradients of
        this batch in self.gp.grads and returns what the discriminator phase needs"""
        Gpp, Gpr, Dpost, Dpre = self.Gpp, self.Gpr, self.Dpost, self.Dpre
        B, C, H, W = x_pre.shape
        n_cond = C - 3
        lb = self.loss_buf
        gpp0, gpr0 = self.gp.grads_of
        (gflat1, (gpp1, gpr1)) = self.g_extra[0]
        dpost_g, dpre_g = self.dp.grads_of

        def cond_into(buf):
            """channels 3.. = the topography conditions input_stack[:, 3:] (model.py:681-689); other channels zero"""
            if n_cond:
                ops.pack_nchw(x_pre[:, 3:], buf, 3, zero_rest=True)

        def packed_input(rgb):
            """generator input [rgb | conditions] as a packed, reflect-haloed operand"""
            buf = ActBuf(B, H, W, 16, halo=3, zero=False)
            if n_cond:
                cond_into(buf)
                ops.pack_nchw(rgb, buf, 0)
            else:
                ops.pack_nchw(rgb, buf, 0, zero_rest=True)
            return buf

        # discriminator inputs: [real | history] halves, and the current synthetic images for the generator phase
        din_pre = ActBuf(2 * B, H, W, 16, zero=False)
        din_post = ActBuf(2 * B, H, W, 16, zero=False)
        ops.pack_nchw(x_pre, din_pre.batch_slice(0, B), 0, zero_rest=True)              # real pre = the input stack
        real_post = din_post.batch_slice(0, B)
        if n_cond:
            cond_into(real_post)
            ops.pack_nchw(y_post, real_post, 0)
        else:
            ops.pack_nchw(y_post, real_post, 0, zero_rest=True)
        fake_post = ActBuf(B, H, W, 16, zero=False)
        fake_pre = ActBuf(B, H, W, 16, zero=False)
        if n_cond:
            cond_into(fake_post)
            cond_into(fake_pre)
        else:
            fake_post.t.zero_()
            fake_pre.t.zero_()

        # ---- generator passes (:682-689)
        synth_post, tA = Gpp.forward(x_pre, d_input=fake_post, d_c0=0)
        synth_pre, tB = Gpr.forward(None, d_input=fake_pre, d_c0=0, xin=packed_input(y_post))
        rec_post, tC = Gpp.forward(None, xin=packed_input(synth_pre))
        rec_pre, tD = Gpr.forward(None, xin=packed_input(synth_post))

        # ---- generator update (:692-713)
        if self.identity:                                                                 # :700-702
            (gflat2, (gpp2, gpr2)) = self.g_extra[1]
            idt_post, tE = Gpp.forward(None, xin=packed_input(y_post))
            d_idt = torch.empty_like(idt_post)
            ops.l1_loss(idt_post, y_post, 5.0, 1.0, lb[8:9], dpred=d_idt)
            Gpp.backward(tE, gpp2, dout_nchw=d_idt)
            del tE
            idt_pre, tF = Gpr.forward(x_pre)
            d_idt2 = torch.empty_like(idt_pre)
            ops.l1_loss(idt_pre, x_pre[:, :3], 5.0, 1.0, lb[9:10], dpred=d_idt2)
            Gpr.backward(tF, gpr2, dout_nchw=d_idt2)
            del tF
        # cycle terms (:710-711) through the second passes, back to the synthetic images
        d_rec_post = torch.empty_like(rec_post)
        ops.l1_loss(rec_post, y_post, 10.0, 1.0, lb[3:4], dpred=d_rec_post)
        g_synth_pre = Gpp.backward(tC, gpp1, dout_nchw=d_rec_post, need_dx=True, rgb_only=True)
        del tC
        d_rec_pre = torch.empty_like(rec_pre)
        ops.l1_loss(rec_pre, x_pre[:, :3], 10.0, 1.0, lb[2:3], dpred=d_rec_pre)
        g_synth_post = Gpr.backward(tD, gpr1, dout_nchw=d_rec_pre, need_dx=True, rgb_only=True)
        del tD
        # adversarial terms (:704-709) through the frozen discriminators: only the input gradient
        for D, fake, slot, g_synth in ((Dpost, fake_post, 0, g_synth_post), (Dpre, fake_pre, 1, g_synth_pre)):
            logits, tape = D.forward_buf(fake)
            dlog = ActBuf(B, logits.h, logits.w, 16, zero=False)
            ops.mse_const_loss(logits, 1.0, 1.0, 1.0, lb[slot:slot + 1], dlog)
            d_din = D.backward(tape, dlog, None, need_dx=True)
            ops.unpack_nchw(d_din, g_synth, 0, accumulate=True)  # + the image channels of the discriminator's input gradient
        Gpp.backward(tA, gpp0, dout_nchw=g_synth_post)
        del tA
        Gpr.backward(tB, gpr0, dout_nchw=g_synth_pre)
        del tB
        ops.add_f32(self.gp.grads.flat, gflat1)
        if self.identity:
            ops.add_f32(self.gp.grads.flat, gflat2)
        return synth_post, synth_pre, fake_pre, fake_post, din_pre, din_post

    def _phase_d(self, state):
        """discriminator update terms (:716-735): real images vs images from the history buffers; leaves the
        discriminator gradients of this batch in self.dp.grads"""
        Dpost, Dpre = self.Dpost, self.Dpre
        _, _, fake_pre, fake_post, din_pre, din_post = state
        B = fake_pre.n
        lb = self.loss_buf
        dpost_g, dpre_g = self.dp.grads_of
        ops.history_exchange(fake_pre.t, self._pool("pre", fake_pre), self.hist_ctrl[0:2], din_pre.t[B:])
        ops.history_exchange(fake_post.t, self._pool("post", fake_post), self.hist_ctrl[2:4], din_post.t[B:])
        for D, din, grads, k_real, k_syn in ((Dpre, din_pre, dpre_g, 4, 6), (Dpost, din_post, dpost_g, 5, 7)):
            logits, tape = D.forward_buf(din)
            dlog = ActBuf(2 * B, logits.h, logits.w, 16, zero=False)
            ops.mse_const_loss(logits.batch_slice(0, B), 1.0, 1.0, 0.5, lb[k_real:k_real + 1], dlog.batch_slice(0, B))
            ops.mse_const_loss(logits.batch_slice(B, B), 0.0, 1.0, 0.5, lb[k_syn:k_syn + 1], dlog.batch_slice(B, B))
            D.backward(tape, dlog, grads, need_dx=False)

    def losses(self):
        """Host copy of the last step's losses (one device->host sync), averaged over ranks."""
        buf = self.loss_buf.clone()
        if self.world_size > 1:
            dist.all_reduce(buf, group=self.group)
            buf /= self.world_size
        return dict(zip(self.loss_keys, buf.tolist()))
