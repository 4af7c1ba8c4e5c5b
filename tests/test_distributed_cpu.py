"""World-size-2 `gloo` tests (CPU) of the host-side data-parallel logic: batch sharding of the synthetic loader,
parameter flattening and the bucketed gradient all-reduce driven by grad-ready callbacks (fpgan/trainer.py).
The kernels themselves need a GPU; everything tested here is plain torch.distributed plumbing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _TinyNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Conv2d(4, 8, 3)
        self.b = nn.Conv2d(8, 8, 3)
        self.c = nn.ConvTranspose2d(8, 4, 3)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fpgan.trainer import FlatParams, _BucketReducer
        from models.data import SyntheticLoader
        torch.manual_seed(0)
        net = _TinyNet()
        ref_state = {k: v.clone() for k, v in net.state_dict().items()}
        fp = FlatParams(net)
        # flattening keeps values, names and state_dict semantics
        for k, v in net.state_dict().items():
            assert torch.equal(v, ref_state[k])
        assert fp.flat.numel() == sum(p.numel() for p in net.parameters())
        fp.flat.mul_(2.0)  # parameters are views of the flat buffer
        assert torch.equal(net.a.weight, ref_state["a.weight"] * 2)

        red = _BucketReducer(fp, bucket_bytes=1024)  # tiny buckets -> several of them
        assert len(red.bounds) > 1 and red.bounds[0][0] == 0 and red.bounds[-1][1] == fp.flat.numel()
        for (s0, e0), (s1, e1) in zip(red.bounds[:-1], red.bounds[1:]):
            assert e0 == s1
        fp.grads.flat.copy_(torch.arange(fp.flat.numel(), dtype=torch.float32) * (rank + 1))
        red.start()
        for name, _ in reversed(fp.named):  # backward order; "c.bias" style names are signalled like the executor
            red.ready(name)
        red.finish()
        want = torch.arange(fp.flat.numel(), dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(fp.grads.flat, want), "bucketed all-reduce must equal the sum over ranks"
        # the default: one bucket = one all-reduce once every layer is done (FPG_DDP_BUCKET_MB=0)
        one = _BucketReducer(fp)
        assert one.bounds == [(0, fp.flat.numel())]
        fp.grads.flat.copy_(torch.arange(fp.flat.numel(), dtype=torch.float32) * (rank + 1))
        one.start()
        for name, _ in reversed(fp.named):
            one.ready(name)
        assert len(one.works) == 1
        one.finish()
        assert torch.equal(fp.grads.flat, want)
        # a parameter whose gradient is never signalled (bias before an InstanceNorm) is still reduced by finish()
        fp.grads.flat.fill_(float(rank + 1))
        red.start()
        for name, _ in fp.named:
            if not name.endswith(".bias"):
                red.ready(name)
        red.finish()
        assert torch.all(fp.grads.flat == sum(r + 1 for r in range(world)))

        # loader sharding: the ranks' shards tile the single-process global batch
        full = next(iter(SyntheticLoader(steps=1, batch=4, channels=9, size=8, pin=False)))
        mine = next(iter(SyntheticLoader(steps=1, batch=2, channels=9, size=8, rank=rank, world_size=world, pin=False)))
        assert torch.equal(mine[0], full[0][2 * rank:2 * rank + 2]) and torch.equal(mine[1], full[1][2 * rank:2 * rank + 2])
        results[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        results[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_sharding_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)


def test_binary_metrics_from_counts_match_definitions():
    """torchmetrics Binary* restated from integer confusion counts (reference model.py:409-418)."""
    import importlib
    M = importlib.import_module("models.model")  # conftest.py puts the package directory on sys.path
    torch.manual_seed(0)
    pred, truth = (torch.rand(5000) > 0.6).float(), (torch.rand(5000) > 0.5).float()
    p, t = pred > 0.5, truth > 0.5
    tp, fp, tn, fn = int((p & t).sum()), int((p & ~t).sum()), int((~p & ~t).sum()), int((~p & t).sum())
    m = M.Model.binary_metrics_from_counts(tp, fp, tn, fn)
    assert abs(m["MSE"] - torch.mean((pred - truth) ** 2).item()) < 1e-7
    assert abs(m["Accuracy"] - (pred == truth).float().mean().item()) < 1e-7
    prec, rec = tp / (tp + fp), tp / (tp + fn)
    assert abs(m["F1_Flood"] - 2 * prec * rec / (prec + rec)) < 1e-12
    inv_p, inv_t = torch.abs(pred - 1) > 0.5, torch.abs(truth - 1) > 0.5
    assert abs(m["Precision_No_Flood"] - int((inv_p & inv_t).sum()) / int(inv_p.sum())) < 1e-12
    assert abs(m["Recall_No_Flood"] - int((inv_p & inv_t).sum()) / int(inv_t.sum())) < 1e-12
    assert M.Model.binary_metrics_from_counts(0, 0, 10, 0)["Precision_Flood"] == 0.0  # safe division


def test_peer_exchange_bucket_plan():
    """fpgan/peer.py::plan_buckets: buckets are contiguous, cover the flat buffer exactly once, every name maps to the
    bucket that holds it, the tail is the small prefix (first layers = last gradients), the others reach the target"""
    from fpgan import peer
    sizes, off = [], 0
    layers = [("stem", 28224, 64), ("c2", 73728, 128), ("c3", 294912, 256)]
    layers += [(f"b{i}.c{j}", 589824, 256) for i in range(9) for j in (1, 2)]
    layers += [("d1", 294912, 128), ("d2", 73728, 64), ("head", 84672, 27)]
    for n, w, b in layers:
        sizes.append((n + ".weight", off, w))
        off += w
        sizes.append((n + ".bias", off, b))
        off += b
    for bucket_mb, tail_mb in ((4, 2), (8, 2), (1, 0.05), (64, 64), (0.001, 0.001)):
        bounds, of, tail = peer.plan_buckets(sizes, int(bucket_mb * 2 ** 20), int(tail_mb * 2 ** 20))
        assert bounds[0][0] == 0 and bounds[-1][1] == off
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])) and all(e > s for s, e in bounds)
        for n, o, k in sizes:
            s, e = bounds[of[n]]
            assert s <= o and o + k <= e
        assert tail == 0
        if len(bounds) > 1:
            if sizes[0][2] * 4 <= tail_mb * 2 ** 20:  # else no prefix fits and the first bucket in flat order is the tail
                assert (bounds[0][1] - bounds[0][0]) * 4 <= tail_mb * 2 ** 20
            for s, e in bounds[2:]:  # every bucket but the tail and its neighbour reaches the target
                assert (e - s) * 4 >= min(bucket_mb * 2 ** 20, max(k for _, _, k in sizes) * 4) * 0.5
    bounds, of, tail = peer.plan_buckets(sizes, 4 << 20, 2 << 20)
    assert of["stem.weight"] == of["c3.bias"] == 0 and of["b0.c1.weight"] != 0
    assert (bounds[0][1] - bounds[0][0]) == sum(w + b for n, w, b in layers[:3])
