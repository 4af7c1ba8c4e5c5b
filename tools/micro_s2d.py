"""Times the PatchGAN stem both ways (CUDA events, L2 flushed): strided 4x4 kernel on the 16-channel input vs
space-to-depth copy + 2x2 stride-1 kernel. Usage: python tools/micro_s2d.py [batch]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from fpgan import ops  # noqa: E402
from models import model_architectures as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
D = A.PairedAttentionDiscriminator(9).cuda()
ex = D._executor()
ex.repack()
din = ops.ActBuf(B, 256, 256, 16, zero=False)
din.t.normal_()
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
    return ts[len(ts) // 2] * 1000


xs = ops.ActBuf(B, 129, 129, 64, zero=False)
print(f"B={B} strided 4x4 on 16 channels : {timed(lambda: ex._conv(din, 'model.0', act=ops.ACT_LEAKY)):7.1f} us")
print(f"B={B} space-to-depth copy        : {timed(lambda: ops.space_to_depth16(din, xs)):7.1f} us")
print(f"B={B} 2x2 on the copy            : {timed(lambda: ex._stem_conv(din, 'model.0', ops.ACT_LEAKY, xs)):7.1f} us")
