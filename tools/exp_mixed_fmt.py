"""Experiment: does tcgen05.mma kind::f16 accept A = bf16 with B = fp16 (mixed operand formats)?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flood-prediction-gan_b200"))
import torch, torch.nn.functional as F
from fpgan import ops

torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
n, c, k, h = 2, 64, 128, 32
x = torch.randn(n, c, h, h, device="cuda", generator=g)
wt = (torch.randn(k, c, 3, 3, device="cuda", generator=g) / 24).requires_grad_(True)
xh = x.half().float()       # activation stored as fp16
y = F.conv2d(xh, wt, None, padding=1)
dy = torch.randn_like(y).to(torch.bfloat16).float()
(ref,) = torch.autograd.grad(y, wt, dy)
spec = ops.ConvSpec(3, 3, 1, 1, c, k)
xb = ops.ActBuf(n, h, h, c)
xb.t = xb.t.view(torch.float16); xb.t.copy_(xh.permute(0, 2, 3, 1))   # same storage, fp16 payload
dyb = ops.ActBuf.from_nchw(dy)
dw = torch.zeros_like(ref)
os.environ["FPG_EXPERIMENT_IDESC_XOR"] = hex(1 << 10)   # B (= Y operand = input x) format bf16 -> f16
try:
    ops.conv_wgrad(xb, dyb, spec, dw)
    torch.cuda.synchronize()
    err = ((dw - ref).norm() / ref.norm()).item()
    print("wgrad mixed bf16 x fp16: rel err", err)
except Exception as e:
    print("wgrad mixed FAILED:", e)
# fprop: A = activation fp16, B = weights bf16 -> flip A format (bit 7)
os.environ["FPG_EXPERIMENT_IDESC_XOR"] = hex(1 << 7)
wb = wt.detach().to(torch.bfloat16).float()
spec.pack(wb.contiguous())
yref = F.conv2d(xh, wb, None, padding=1)
yb = ops.ActBuf(n, h, h, k, fp32=True)
try:
    ops.conv_fprop(xb, spec, yb)
    torch.cuda.synchronize()
    print("fprop mixed fp16 x bf16: rel err", ((yb.to_nchw() - yref).norm() / yref.norm()).item())
except Exception as e:
    print("fprop mixed FAILED:", e)
