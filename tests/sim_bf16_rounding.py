"""CPU simulation of where bf16 rounding enters the generator forward (weights, conv inputs, conv outputs,
residual stream) using the oracle's graph. Shows that the measured 2.25e-2 output deviation of the kernels is the
bf16 floor, and what each rounding site contributes. Not part of the product (it lives under tests/ because it executes the oracle)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import gan_oracle as O  # noqa: E402


def r(t, on=True):
    return t.to(torch.bfloat16).float() if on else t


def forward(p, x, rw=True, rz=True, ry=True, rs=True):
    def conv(t, name, **kw):
        return r(F.conv2d(r(t, rz), r(p[name + ".weight"], rw), None, **kw), ry)

    def conv_t(t, name):
        return r(F.conv_transpose2d(r(t, rz), r(p[name + ".weight"], rw), None, stride=2, padding=1,
                                    output_padding=1), ry)

    def inorm(t):
        return F.instance_norm(t, eps=1e-5)

    t = F.relu(inorm(conv(F.pad(x, (3,) * 4, "reflect"), "conv1")))
    t = F.relu(inorm(conv(t, "conv2", stride=2, padding=1)))
    t = r(F.relu(inorm(conv(t, "conv3", stride=2, padding=1))), rs)
    for i in range(9):
        b = f"resnet_blocks.{i}."
        y = F.relu(inorm(conv(F.pad(t, (1,) * 4, "reflect"), b + "conv1")))
        y = inorm(conv(F.pad(y, (1,) * 4, "reflect"), b + "conv2"))
        t = r(t + y, rs)

    def up(t, n):
        return F.relu(inorm(conv_t(t, n)))

    c = F.pad(up(up(t, "deconv1_content"), "deconv2_content"), (3,) * 4, "reflect")
    content = torch.tanh(F.conv2d(r(c, rz), r(p["deconv3_content.weight"], rw), p["deconv3_content.bias"]))
    a = up(up(t, "deconv1_attention"), "deconv2_attention")
    att = torch.softmax(F.conv2d(r(a, rz), r(p["deconv3_attention.weight"], rw), p["deconv3_attention.bias"]), dim=1)
    out = content[:, 0:3] * att[:, 0:1]
    for k in range(1, 9):
        out = out + content[:, 3 * k:3 * k + 3] * att[:, k:k + 1]
    return out + x[:, :3] * att[:, 9:10]


if __name__ == "__main__":
    p = O.init_model("pairedattention", "all", 47)["generator"]
    x, _ = O.synthetic_batch(0, 2, 9, 64)
    with torch.no_grad():
        ref = forward(p, x, False, False, False, False)
        for name, kw in [("all four sites", {}), ("weights only", dict(rz=False, ry=False, rs=False)),
                         ("conv inputs only", dict(rw=False, ry=False, rs=False)),
                         ("conv outputs only", dict(rw=False, rz=False, rs=False)),
                         ("residual stream only", dict(rw=False, rz=False, ry=False)),
                         ("without conv-output and stream rounding", dict(ry=False, rs=False))]:
            o = forward(p, x, **kw)
            print(f"{name:42s} rel-rms err {((o - ref).norm() / ref.norm()).item():.4f}")
