"""Experiment: how much of the step is launch gaps / host overhead? Captures ONE fused PairedAttention step into a CUDA
graph (Adam step count frozen: timing only, not a training loop) and compares replay time with eager stepping."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fpgan.trainer import PairedTrainer  # noqa: E402
from models import model_architectures as A  # noqa: E402

B = 16
torch.manual_seed(47)
G, D = A.PairedAttentionGenerator(9).cuda(), A.PairedAttentionDiscriminator(9).cuda()
tr = PairedTrainer(G, D)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(B, 9, 256, 256, device="cuda", generator=g) * 2 - 1
y = torch.rand(B, 3, 256, 256, device="cuda", generator=g) * 2 - 1
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()


def timed(fn, reps=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


eager = timed(lambda: tr.step(x, y))
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    tr.step(x, y)
torch.cuda.synchronize()
graph.replay()
torch.cuda.synchronize()
replay = timed(graph.replay)
print(f"eager {eager:.3f} ms/step, graph replay {replay:.3f} ms/step ({100 * (eager - replay) / eager:.1f}% less)")
