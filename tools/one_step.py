"""Runs warm-up steps of the fused PairedAttention step, then ONE step between cudaProfilerStart/Stop
(for `ncu --profile-from-start off`). Usage: python tools/one_step.py [batch] [size]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fpgan.trainer import PairedTrainer  # noqa: E402
from models import model_architectures as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(47)
G, D = A.PairedAttentionGenerator(9).cuda(), A.PairedAttentionDiscriminator(9).cuda()
tr = PairedTrainer(G, D)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(B, 9, S, S, device="cuda", generator=g) * 2 - 1
y = torch.rand(B, 3, S, S, device="cuda", generator=g) * 2 - 1
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", tr.losses())
