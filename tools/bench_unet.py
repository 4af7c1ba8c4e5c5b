"""BASELINE.json configs[4]: segmentation U-Net inference on generated vs ground-truth 256x256 tiles, batch 64
(reference model.py:380-418): two U-Net forward passes per tile pair, flood-mask threshold, confusion counts.
Prints tile pairs/s and algorithmic TFLOP/s (96.33 GFLOP per forward)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from fpgan import ops  # noqa: E402
from models import model_architectures as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(47)
net = A.UNet().cuda()
g = torch.Generator(device="cuda").manual_seed(0)
gen = torch.rand(B, 3, S, S, device="cuda", generator=g) * 2 - 1
truth = torch.rand(B, 3, S, S, device="cuda", generator=g) * 2 - 1
for _ in range(2):
    A.flood_masks_and_counts(net, gen, truth)
torch.cuda.synchronize()
ops.PROFILE = {}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
a.record()
for _ in range(steps):
    _, _, counts = A.flood_masks_and_counts(net, gen, truth)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
prof = {k: sum(x.elapsed_time(y) for x, y in v) / steps for k, v in ops.PROFILE.items()}
ops.PROFILE = None
gflop = 96.33 * (S / 256) ** 2 * 2 * B
print(f"U-Net inference B={B} {S}x{S}: {ms:.2f} ms per batch of tile pairs = {B / ms * 1000:.0f} pairs/s, "
      f"{gflop / ms:.1f} TFLOP/s algorithmic; counts {counts.tolist()}")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:10]:
    print(f"   {v:8.3f} ms  {k}")
