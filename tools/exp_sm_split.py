"""Experiment: two half-batches on two streams, every grid capped at half the SMs (FPG_SM_BUDGET), against one full
batch on the whole chip -- does one half's convolutions overlap the other half's normalisation passes?
Usage: FPG_SM_BUDGET=74 python tools/exp_sm_split.py 8 2     (two trainers of batch 8)
       python tools/exp_sm_split.py 16 1                     (baseline)"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from fpgan.trainer import PairedTrainer  # noqa: E402
from models import model_architectures as A  # noqa: E402

B, N = int(sys.argv[1]), int(sys.argv[2])
steps = 30
trainers, streams, data = [], [], []
for i in range(N):
    torch.manual_seed(47 + i)
    trainers.append(PairedTrainer(A.PairedAttentionGenerator(9).cuda(), A.PairedAttentionDiscriminator(9).cuda()))
    streams.append(torch.cuda.Stream())
    g = torch.Generator(device="cuda").manual_seed(i)
    data.append((torch.rand(B, 9, 256, 256, device="cuda", generator=g) * 2 - 1,
                 torch.rand(B, 3, 256, 256, device="cuda", generator=g) * 2 - 1))
torch.cuda.synchronize()


def one_round():
    for tr, st, (x, y) in zip(trainers, streams, data):
        with torch.cuda.stream(st):
            tr.step(x, y)


for _ in range(5):  # eager calls, graph capture, first replays
    one_round()
    torch.cuda.synchronize()
offset = int(os.environ.get("FPG_EXP_OFFSET_CYCLES", "0"))  # phase offset of the second stream (lockstep breaker)
if offset and N > 1:
    with torch.cuda.stream(streams[1]):
        torch.cuda._sleep(offset)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    one_round()
for st in streams:
    torch.cuda.current_stream().wait_stream(st)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(f"{N} trainer(s) x batch {B}, SM budget {os.environ.get('FPG_SM_BUDGET', 'all')}: {ms:.3f} ms per round = "
      f"{N * B * 1000 / ms:.0f} tiles/s")
