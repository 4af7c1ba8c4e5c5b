"""Micro-benchmark of the InstanceNorm kernels on the residual-trunk shape (for ncu / event timing)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from fpgan import ops  # noqa: E402

n, h, w, c = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 64, 64, 256)))
halo = int(sys.argv[5]) if len(sys.argv) > 5 else 1
y = ops.ActBuf(n, h, w, c)
y.t.normal_()
res = ops.ActBuf(n, h, w, c, halo=halo)
res.t.normal_()
z = ops.ActBuf(n, h, w, c, halo=halo)
dz = ops.ActBuf(n, h, w, c, halo=halo)
dz.t.normal_()
dz2 = ops.ActBuf(n, h, w, c)
dz2.t.normal_()
dy, dres = ops.ActBuf(n, h, w, c), ops.ActBuf(n, h, w, c)
stats = torch.empty(n * c * 2, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


elems = n * h * w * c
for name, fn, bytes_ in (
        ("stats", lambda: ops.instnorm_stats(y, stats), elems * 2),
        ("apply+res+halo", lambda: ops.instnorm_apply(y, stats, 1, z, residual=res), elems * 6),
        ("bwd fold+dz2+dres", lambda: ops.instnorm_bwd(dz, y, stats, 0, dy, dz2=dz2, dres=dres), elems * 14),
        ("bwd fold", lambda: ops.instnorm_bwd(dz, y, stats, 1, dy), elems * 10)):
    ms = timed(fn)
    print(f"{name:20s} {ms * 1000:8.1f} us  {bytes_ / ms / 1e6:8.1f} GB/s (algorithmic bytes {bytes_ / 1e6:.1f} MB)")
