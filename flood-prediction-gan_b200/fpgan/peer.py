"""Data-parallel gradient exchange through peer memory (csrc/peer.cu; C ABI "Peer exchange" in include/fpg.h).

What stands between `loss.backward()` and `optimizer.step()` (models/model.py:632-633, :645-646) when the batch is
sharded over W GPUs of one NVSwitch box. Instead of an all-reduce done by kernels (NCCL), every rank PUSHES its gradient
buckets into a staging slot on every peer with the copy engines -- no SM is taken from the persistent convolution grids,
so the transfers overlap the backward pass for free -- and the Adam kernel sums the W sources in rank order
(`fpg_adam_step_dev_multi`): replicas stay bit-identical and equal one process accumulating the shards in shard order
(`PairedTrainer.step_accumulated`). Cross-rank ordering = flag words written into peer memory (release stores at system
scope, polled locally), with values taken from a device-side step counter: every launch has constant arguments, the
step replays as a CUDA graph.

`PeerReducer` has the interface of `trainer._BucketReducer` (start / ready / finish) plus `adam()`.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import lib as L
from . import ops

_FLAG_BYTES = 256      # per rank: data flags u32[16] at +0, ack flags u32[16] at +64
_ACK_OFF = 64
MAX_WORLD = 16


def supported(group=None):
    """True when every rank of the group sits in this process's node and can map the others' memory (NVLink / PCIe
    peer access); evaluated collectively so that all ranks take the same path."""
    if not (dist.is_available() and dist.is_initialized()):
        return False
    world = dist.get_world_size(group)
    if world < 2 or world > MAX_WORLD or not torch.cuda.is_available():
        return False
    ok = os.environ.get("FPG_DDP", "auto") != "nccl" and torch.cuda.device_count() >= world
    dev = torch.cuda.current_device()
    if ok:
        # one process per GPU of ONE node (torchrun --nnodes=1): the ranks own devices 0..W-1 of this node
        ok = all(o == dev or torch.cuda.can_device_access_peer(dev, o) for o in range(world))
    names = [None] * world
    dist.all_gather_object(names, (os.uname().nodename, dev, bool(ok)), group=group)
    same_node = len({n[0] for n in names}) == 1
    distinct = len({n[1] for n in names}) == world
    return same_node and distinct and all(n[2] for n in names)


def plan_buckets(named_sizes, bucket_bytes, tail_bytes):
    """Buckets over a flat parameter buffer laid out in `named_sizes` order [(name, offset, numel)] (forward order of
    the network) for a backward pass that produces the gradients roughly back to front. Returns (bounds [(start, end)]
    in elements, bucket index per name, tail bucket index):
      * the TAIL bucket = the longest prefix of at most tail_bytes (the first layers: their gradients complete when
        the backward pass ends, so their transfer cannot be hidden -- keep it small; it is written by one kernel);
      * the rest is cut from the END of the buffer into contiguous buckets of at least bucket_bytes (a bucket closes
        at the first layer boundary at or past the target), so that the buckets that complete first are full-sized.
    Pure host logic (tested on CPU)."""
    n = len(named_sizes)
    total_end = named_sizes[-1][1] + named_sizes[-1][2] if n else 0
    n_tail = 0
    while n_tail < n - 1 and (named_sizes[n_tail][1] + named_sizes[n_tail][2]) * 4 <= tail_bytes:
        n_tail += 1
    tail_end = named_sizes[n_tail][1] if n_tail < n else total_end
    groups = []  # from the end: [first index, last index]
    i = n - 1
    while i >= n_tail:
        j, end = i, named_sizes[i][1] + named_sizes[i][2]
        while j > n_tail and (end - named_sizes[j][1]) * 4 < bucket_bytes:
            j -= 1
        groups.append((j, i))
        i = j - 1
    # a short remainder next to the tail joins its neighbour instead of becoming a bucket of its own
    if len(groups) >= 2:
        j, i2 = groups[-1]
        if (named_sizes[i2][1] + named_sizes[i2][2] - named_sizes[j][1]) * 4 < bucket_bytes // 2:
            pj, pi = groups[-2]
            groups[-2:] = [(j, pi)]
    bounds, of = [], {}
    if n_tail > 0:
        bounds.append((0, tail_end))
        for k in range(n_tail):
            of[named_sizes[k][0]] = 0
    tail = 0 if n_tail > 0 else None
    for j, i2 in reversed(groups):
        b = len(bounds)
        bounds.append((named_sizes[j][1], named_sizes[i2][1] + named_sizes[i2][2]))
        for k in range(j, i2 + 1):
            of[named_sizes[k][0]] = b
    if tail is None:  # everything fits no prefix: the first bucket in flat order completes last
        tail = 0
    return bounds, of, tail


class PeerReducer:
    """Gradient exchange of one FlatParams over peer memory. Per step: start() -> ready(name)... -> finish() -> adam()."""

    def __init__(self, flat_params, group=None, bucket_bytes=None, keep_sum=None):
        self.fp = flat_params
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert 2 <= self.world <= MAX_WORLD
        if bucket_bytes is None:
            # a copy-engine copy costs ~14 us + bytes / 620 GB/s (tools/micro_peer.py) and a rank issues W-1 of them per
            # bucket: 4 MB buckets = 7 x 21 us per residual block of the generator, well under the block's backward time
            bucket_bytes = int(float(os.environ.get("FPG_PEER_BUCKET_MB", "4")) * (1 << 20))
        tail_bytes = int(float(os.environ.get("FPG_PEER_TAIL_MB", "2")) * (1 << 20))
        if keep_sum is None:  # write the summed gradient back into grads.flat (checks that compare gradients)
            keep_sum = os.environ.get("FPG_PEER_KEEP_SUM", "0") == "1"
        self.keep_sum = keep_sum
        self.timeout_s = float(os.environ.get("FPG_PEER_TIMEOUT_S", "120"))
        sizes = [(n,) + tuple(flat_params.offsets[n]) for n, _ in flat_params.named]
        self.bounds, self.layer_bucket, self.tail = plan_buckets(sizes, bucket_bytes, tail_bytes)
        self.count = flat_params.flat.numel()
        self.slot_bytes = (self.count * 4 + 255) // 256 * 256
        self.flags_off = self.world * self.slot_bytes
        total = self.flags_off + _FLAG_BYTES
        dev = flat_params.flat.device
        self.lib = L.load()
        base = C.c_void_p()
        L.call("fpg_peer_alloc", C.byref(base), total)
        self.base = base.value
        handle = (C.c_ubyte * 64)()
        L.call("fpg_peer_export", C.c_void_p(self.base), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_base = [None] * self.world
        for p in range(self.world):
            if p == self.rank:
                self.peer_base[p] = self.base
                continue
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[p])
            ptr = C.c_void_p()
            L.call("fpg_peer_open", buf, C.byref(ptr))
            self.peer_base[p] = ptr.value
        # peers in ring order starting after this rank, so that at any moment the W ranks address W different targets
        self.peers = [(self.rank + i) % self.world for i in range(1, self.world)]
        self.ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.status = torch.zeros(2, dtype=torch.int64, device=dev)  # {lane that timed out + 1, ns spent waiting}
        n_streams = max(1, int(os.environ.get("FPG_PEER_STREAMS", "4")))
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(min(n_streams, len(self.peers)))]
        self._data_ptrs = (C.c_void_p * len(self.peers))(
            *[self.peer_base[p] + self.flags_off + 4 * self.rank for p in self.peers])
        self._ack_ptrs = (C.c_void_p * len(self.peers))(
            *[self.peer_base[p] + self.flags_off + _ACK_OFF + 4 * self.rank for p in self.peers])
        self._sources = (C.c_void_p * self.world)(
            *[self.fp.grads.flat.data_ptr() if p == self.rank else self.base + p * self.slot_bytes
              for p in range(self.world)])
        self.pending = None
        self._pushed = None
        self._forked = False
        self.overlap = os.environ.get("FPG_PEER_OVERLAP", "1") != "0"  # 0: everything after the backward pass
        self.kernel_push_max = int(float(os.environ.get("FPG_PEER_KERNEL_PUSH_MB", "8")) * (1 << 20))
        self.closed = False
        dist.barrier(group=group)  # every rank has mapped every block before the first push

    # ---- per step
    def start(self):
        """before the backward pass: the peers have consumed the previous step's staged gradients (their Adam ran at
        about the time ours did, so this never waits in practice)"""
        ops._run("peer_wait", 1, "fpg_peer_wait", C.c_void_p(self.base + self.flags_off + _ACK_OFF), self.world,
                 self.rank, ops._ptr(self.ctr), 0, ops._ptr(self.status), self.timeout_s, ops._stream())
        self.pending = [0] * len(self.bounds)
        for n in self.layer_bucket:
            self.pending[self.layer_bucket[n]] += 1
        self._pushed = [False] * len(self.bounds)

    def _push(self, b):
        """bucket b -> its slot on every peer, by the copy engines, behind the stream that produced the gradients"""
        s, e = self.bounds[b]
        self._pushed[b] = True
        if e <= s:
            return
        ev = torch.cuda.Event()
        ev.record()
        src = self.fp.grads.flat.data_ptr() + 4 * s
        for st in self.streams:
            st.wait_event(ev)
        self._forked = True
        for i, p in enumerate(self.peers):
            st = self.streams[i % len(self.streams)]
            dst = self.peer_base[p] + self.rank * self.slot_bytes + 4 * s
            L.call("fpg_peer_copy", C.c_void_p(dst), C.c_void_p(src), 4 * (e - s), C.c_void_p(st.cuda_stream))

    def _push_tail(self, b):
        """bucket b by ONE kernel on the current stream (the backward pass is over: nothing to take SMs from)"""
        s, e = self.bounds[b]
        self._pushed[b] = True
        lo, hi = s // 4 * 4, (e + 3) // 4 * 4  # 16-byte granules (the slots are padded; neighbours hold the same data)
        hi = min(hi, self.slot_bytes // 4)
        if hi <= lo:
            return
        dsts = (C.c_void_p * len(self.peers))(
            *[self.peer_base[p] + self.rank * self.slot_bytes + 4 * lo for p in self.peers])
        ops._run("peer_push", 1, "fpg_peer_push", C.c_void_p(self.fp.grads.flat.data_ptr() + 4 * lo), dsts,
                 len(self.peers), 4 * (hi - lo), ops._stream())

    def ready(self, name):
        b = self.layer_bucket[name]
        self.pending[b] -= 1
        if self.pending[b] == 0 and b != self.tail and self.overlap:
            self._push(b)

    def finish(self):
        """send what is left -- the tail bucket and anything whose gradient was never announced (it keeps its zeros)
        -- tell the peers, and make the current stream wait until every peer's gradients have landed here"""
        if self.pending is None:
            self.start()
        left = [b for b in range(len(self.bounds)) if not self._pushed[b]]
        big = sum(self.bounds[b][1] - self.bounds[b][0] for b in left) * 4 > self.kernel_push_max
        for b in left:
            (self._push if big else self._push_tail)(b)
        cur = torch.cuda.current_stream()
        if self._forked:  # join the copy streams (required inside a graph capture) before the flags go out
            for st in self.streams:
                ev = torch.cuda.Event()
                ev.record(st)
                cur.wait_event(ev)
            self._forked = False
        ops._run("peer_signal", 1, "fpg_peer_signal", self._data_ptrs, len(self.peers), ops._ptr(self.ctr), 1, 0,
                 ops._stream())
        ops._run("peer_wait", 1, "fpg_peer_wait", C.c_void_p(self.base + self.flags_off), self.world, self.rank,
                 ops._ptr(self.ctr), 1, ops._ptr(self.status), self.timeout_s, ops._stream())
        self.pending = None

    def waited_ms(self):
        """milliseconds this rank's stream has spent in the flag waits so far (device-side clock; one host sync)"""
        return self.status[1].item() / 1e6

    def adam(self, grad_scale, betas=(0.5, 0.999), eps=1e-8, fused=None):
        """Adam on the rank-ordered sum of the W gradient sources (with `fused`, a networks.AdamPack: in the same launch
        as the operand repack), then release the staging slots to the peers"""
        fp = self.fp
        gsum = fp.grads.flat if self.keep_sum else None
        if fused is not None:
            fused.run(grad_scale, sources=self._sources, n_src=self.world, gsum=gsum, betas=betas, eps=eps)
        else:
            ops._run("adam_step", 2, "fpg_adam_step_dev_multi", ops._ptr(fp.flat), self._sources, self.world,
                     ops._ptr(fp.m), ops._ptr(fp.v), fp.flat.numel(), float(betas[0]), float(betas[1]), float(eps),
                     ops._ptr(fp.state), float(grad_scale), ops._ptr(gsum) if gsum is not None else None, ops._stream())
        ops._run("peer_signal", 1, "fpg_peer_signal", self._ack_ptrs, len(self.peers), ops._ptr(self.ctr), 1, 1,
                 ops._stream())

    def close(self):
        """collective: unmap the peers' blocks and free ours (after everybody has stopped using them)"""
        if self.closed:
            return
        self.closed = True
        torch.cuda.synchronize()
        try:
            dist.barrier(group=self.group)
        except Exception:  # the process group is already gone: nothing else can be running either
            pass
        for p in range(self.world):
            if p != self.rank and self.peer_base[p] is not None:
                self.lib.fpg_peer_close(C.c_void_p(self.peer_base[p]))
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        self.lib.fpg_peer_free(C.c_void_p(self.base))
