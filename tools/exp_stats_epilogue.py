import os, sys
sys.path.insert(0, "/root/repo/flood-prediction-gan_b200")
import torch
from fpgan import ops, lib as L
n,h,w,c,k=16,64,64,256,256
spec = ops.ConvSpec(3,3,1,0,c,k); spec.pack((torch.randn(k,c,3,3,device="cuda")*0.02).contiguous())
x = ops.ActBuf(n,h,w,c,halo=1); x.t.normal_()
y = ops.ActBuf(n,h,w,k)
stats = torch.empty(n*k*2, device="cuda")
print("rows", L.load().fpg_conv_stats_rows(x.ref(), spec.gref(), y.ref(), 0))
def timed(fn, reps=20):
    for _ in range(3): fn()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps*1000
print("conv only      %.1f us" % timed(lambda: ops.conv_fprop(x, spec, y)))
print("conv+statskern %.1f us" % timed(lambda: (ops.conv_fprop(x, spec, y), ops.instnorm_stats(y, stats))))
print("conv w/ stats  %.1f us" % timed(lambda: ops.conv_with_stats(x, spec, y, stats)))
ops.PROFILE = {}
ops.conv_with_stats(x, spec, y, stats); torch.cuda.synchronize()
for k_, v in ops.PROFILE.items(): print(k_, [a.elapsed_time(b)*1000 for a,b in v])
