"""Micro-benchmark of the conv kernels on one layer shape (CUDA events, L2 flushed between repetitions).
Usage: python tools/micro_conv.py n h w c k r stride pad halo [fprop|dgrad|wgrad ...]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from fpgan import ops  # noqa: E402

n, h, w, c, k, r, stride, pad, halo = (int(v) for v in sys.argv[1:10])
kinds = sys.argv[10:] or ["fprop", "dgrad", "wgrad"]
cp, kp = ops.pad16(c), ops.pad16(k)
spec = ops.ConvSpec(r, r, stride, pad, cp, kp, c_in_valid=c, c_out_valid=k)
wt = torch.randn(k, c, r, r, device="cuda") * 0.05
spec.pack(wt.contiguous())
hp, wp = h + 2 * halo, w + 2 * halo
ho, wo = (hp + 2 * pad - r) // stride + 1, (wp + 2 * pad - r) // stride + 1
x = ops.ActBuf(n, h, w, cp, halo=halo)
x.t.normal_()
y = ops.ActBuf(n, ho, wo, kp)
dy = ops.ActBuf(n, ho, wo, kp)
dy.t.normal_()
dx = ops.ActBuf(n, h, w, cp, halo=halo)
dw = torch.empty_like(wt)
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


flops = 2.0 * n * ho * wo * k * c * r * r
for kind in kinds:
    fn = {"fprop": lambda: ops.conv_fprop(x, spec, y), "dgrad": lambda: ops.conv_dgrad(dy, spec, dx),
          "wgrad": lambda: ops.conv_wgrad(x, dy, spec, dw)}[kind]
    ms = timed(fn)
    print(f"{kind:6s} n{n} {h}x{w} c{c} k{k} r{r} s{stride}: {ms * 1000:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s "
          f"(algorithmic, unpadded)")
