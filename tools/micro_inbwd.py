"""Micro-benchmark of the data-gradient kernel with the InstanceNorm-backward reductions in its epilogue on the residual
layer shape (n16 64x64 c256): plain dgrad, fused relu-type (one staged operand), fused residual-type with the
skip-gradient merge (three staged operands), and the separate streamed passes they replace.
Usage: python tools/micro_inbwd.py"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from fpgan import ops  # noqa: E402

n, h, c = 16, 64, 256
spec = ops.ConvSpec(3, 3, 1, 0, c, c)
spec.pack((torch.randn(c, c, 3, 3, device="cuda") * 0.05).contiguous())
dy = ops.ActBuf(n, h, h, c)
dy.t.normal_()
y = ops.ActBuf(n, h, h, c, f16=True)
y.t.normal_()
stats = torch.empty(n * c * 2, device="cuda")
ops.instnorm_stats(y, stats)
z = ops.ActBuf(n, h, h, c, halo=1)
z.t.normal_()
zprev = ops.ActBuf(n, h, h, c, halo=1)
zprev.t.normal_()
add = ops.ActBuf(n, h, h, c, halo=1)
add.t.normal_()
dx = ops.ActBuf(n, h, h, c, halo=1, zero=False)
dyo = ops.ActBuf(n, h, h, c, zero=False)
gres = ops.ActBuf(n, h, h, c, zero=False)
dz2 = ops.ActBuf(n, h, h, c)
dz2.t.normal_()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=7):
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


tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("FPG_"))
print(f"[{tag}]")
print("plain dgrad                 %.1f us" % (1e3 * timed(lambda: ops.conv_dgrad(dy, spec, dx))))
print("fused relu-type             %.1f us" % (1e3 * timed(lambda: ops.conv_dgrad_inbwd(dy, spec, dx, z, force=True))))
print("fused residual-type         %.1f us" % (1e3 * timed(lambda: ops.conv_dgrad_inbwd(dy, spec, dx, z, zprev, force=True))))
print("fused residual-type + add   %.1f us" % (1e3 * timed(lambda: ops.conv_dgrad_inbwd(dy, spec, dx, z, zprev, add, force=True))))
red = torch.zeros(n, c, 2, device="cuda")
print("apply pass only             %.1f us" % (1e3 * timed(lambda: ops.instnorm_bwd_apply(dx, y, stats, red, ops.ACT_NONE, dyo))))
print("two-pass IN bwd (relu)      %.1f us" % (1e3 * timed(lambda: ops.instnorm_bwd(dx, y, stats, ops.ACT_RELU, dyo))))
print("two-pass IN bwd (res + add) %.1f us" % (1e3 * timed(lambda: ops.instnorm_bwd(dx, y, stats, ops.ACT_NONE, dyo, dz2=dz2, dres=gres))))
