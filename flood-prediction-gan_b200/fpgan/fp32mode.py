"""fp32 parity mode of the PairedAttention / AttentionGAN generator and the InstanceNorm PatchGAN (forward + losses).

north_star: "Generator outputs and per-step losses must match the reference ... fp32 rtol 1e-4". The reference is fp32
end to end (models/model.py:24, model_architectures.py:339-400); tcgen05 has no fp32 MMA. In this mode every
tensor-core operand is carried as a PAIR of bf16 tensors, v = hi + lo (16 mantissa bits), and a convolution is
evaluated as hi_x*hi_w + lo_x*hi_w + hi_x*lo_w with fp32 accumulation -- as ONE launch of the ordinary implicit-GEMM
kernels: activations hold the channel blocks [hi | lo | hi] (3 * C channels), packed weights the blocks
[hi_w | hi_w | lo_w] along the contraction dimension (see csrc/split.cu). Convolution outputs, InstanceNorm, the
residual stream, tanh / softmax / blend and the losses are fp32. 3x the tensor work and fp32 activations: a mode to
DEMONSTRATE parity of the kernels' arithmetic, not the training path (the bf16 step is bound to 2e-2).

The host code here prepares operands with a few torch calls (weight splitting, torch.cat of the discriminator input);
every convolution, normalisation, blend and loss runs in libfpg_b200.so.
"""
import torch

from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH, ActBuf, ConvSpec, pad16


def split_pad(c):
    """channels of one block of a [hi | lo | hi] operand: multiples of 64, so that 3 blocks tile into the kernels'
    64-channel K chunks (the 9- and 12-channel network inputs are padded to 64)"""
    return -(-c // 64) * 64


class SplitConv:
    """One convolution of the parity mode: weight [K][C][R][S] (nn.Conv2d), or [Cin_T][Cout_T][R][S] (nn.ConvTranspose2d,
    executed as the data gradient of the equivalent forward conv, like the bf16 path)."""

    def __init__(self, weight, bias, r, stride, pad, transposed=False, use_bias=False):
        w = weight.detach().float()
        hi = w.bfloat16().float()
        lo = w - hi
        k, c = w.shape[0], w.shape[1]
        self.transposed = transposed
        if not transposed:
            cp = split_pad(c)
            w3 = torch.zeros(k, 3 * cp, r, r, dtype=torch.float32, device=w.device)
            w3[:, :c], w3[:, cp:cp + c], w3[:, 2 * cp:2 * cp + c] = hi, hi, lo
            self.spec = ConvSpec(r, r, stride, pad, 3 * cp, pad16(k), c_in_valid=3 * cp, c_out_valid=k)
            self.spec.pack(w3.contiguous(), fprop=True, dgrad=False)
            self.c_in, self.c_out, valid = 3 * cp, pad16(k), k
        else:
            kp = split_pad(k)
            w3 = torch.zeros(3 * kp, c, r, r, dtype=torch.float32, device=w.device)
            w3[:k], w3[kp:kp + k], w3[2 * kp:2 * kp + k] = hi, hi, lo
            self.spec = ConvSpec(r, r, stride, pad, pad16(c), 3 * kp, c_in_valid=c, c_out_valid=3 * kp)
            self.spec.pack(w3.contiguous(), fprop=False, dgrad=True)
            self.c_in, self.c_out, valid = 3 * kp, pad16(c), c
        self.bias = None
        if use_bias:
            self.bias = torch.zeros(self.c_out, dtype=torch.float32, device=w.device)
            self.bias[:valid] = bias.detach().float()

    def __call__(self, x3, act=ACT_NONE):
        """x3: [hi | lo | hi] operand -> fp32 NHWC output (bias / activation applied in the conv epilogue)"""
        assert x3.c == self.c_in, (x3.c, self.c_in)
        g = self.spec.g
        if self.transposed:
            y = ActBuf(x3.n, x3.h * 2, x3.w * 2, self.c_out, fp32=True, zero=False)
            ops.conv_dgrad(x3, self.spec, y, bias=self.bias, act=act)
        else:
            hp, wp = x3.h + 2 * x3.halo, x3.w + 2 * x3.halo
            y = ActBuf(x3.n, (hp + 2 * g.pad - g.r) // g.stride + 1, (wp + 2 * g.pad - g.s) // g.stride + 1, self.c_out,
                       fp32=True, zero=False)
            ops.conv_fprop(x3, self.spec, y, bias=self.bias, act=act)
        return y


def _norm(y, act, halo=0, residual=None, want_skip=False, norm=True):
    out = ActBuf(y.n, y.h, y.w, 3 * y.c, halo=halo, zero=False)
    skip = torch.empty(y.n, y.h, y.w, y.c, dtype=torch.float32, device=y.t.device) if want_skip else None
    ops.norm_split_f32(y, out, norm=norm, act=act, residual=residual, skip_out=skip)
    return out, skip


class SplitAttentionGenerator:
    """PairedAttentionGenerator / AttentionGANGenerator forward (model_architectures.py:339-400) in the parity mode."""

    def __init__(self, module):
        m = module
        self.conv1 = SplitConv(m.conv1.weight, None, 7, 1, 0)
        self.conv2 = SplitConv(m.conv2.weight, None, 3, 2, 1)
        self.conv3 = SplitConv(m.conv3.weight, None, 3, 2, 1)
        self.blocks = [(SplitConv(b.conv1.weight, None, 3, 1, 0), SplitConv(b.conv2.weight, None, 3, 1, 0))
                       for b in m.resnet_blocks]
        self.dec = {}
        for br in ("content", "attention"):
            self.dec[br] = (SplitConv(getattr(m, f"deconv1_{br}").weight, None, 3, 2, 1, transposed=True),
                            SplitConv(getattr(m, f"deconv2_{br}").weight, None, 3, 2, 1, transposed=True))
        self.head_content = SplitConv(m.deconv3_content.weight, m.deconv3_content.bias, 7, 1, 0, use_bias=True)
        self.head_attention = SplitConv(m.deconv3_attention.weight, m.deconv3_attention.bias, 1, 1, 0, use_bias=True)

    def forward(self, x):
        """x: fp32 NCHW CUDA [B, C <= 16, H, W] -> (image fp32 [B, 3, H, W], background attention mask [B, H, W])"""
        B, C, H, W = x.shape
        xin = ActBuf(B, H, W, 3 * split_pad(C), halo=3, zero=False)
        ops.pack_nchw_split(x.contiguous(), xin)
        z, _ = _norm(self.conv1(xin), ACT_RELU)
        z, _ = _norm(self.conv2(z), ACT_RELU)
        z, skip = _norm(self.conv3(z), ACT_RELU, halo=1, want_skip=True)
        for i, (c1, c2) in enumerate(self.blocks):
            za, _ = _norm(c1(z), ACT_RELU, halo=1)
            last = i == len(self.blocks) - 1
            z, skip = _norm(c2(za), ACT_NONE, halo=0 if last else 1, residual=skip, want_skip=True)
        heads = {}
        for br, halo in (("content", 3), ("attention", 0)):
            d1, d2 = self.dec[br]
            v, _ = _norm(d1(z), ACT_RELU)
            heads[br], _ = _norm(d2(v), ACT_RELU, halo=halo)
        content = self.head_content(heads["content"], act=ACT_TANH)
        logits = self.head_attention(heads["attention"])
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
        mask = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
        ops.blend_fwd(content, logits, xin, out_nchw=out, mask=mask, input_lo_offset=split_pad(C))
        return out, mask


class SplitPatchGAN:
    """InstanceNorm PatchGAN forward (model_architectures.py:420-441) in the parity mode."""

    def __init__(self, module):
        seq = module.model
        self.c0 = SplitConv(seq[0].weight, seq[0].bias, 4, 2, 1, use_bias=True)
        self.c2 = SplitConv(seq[2].weight, None, 4, 2, 1)
        self.c5 = SplitConv(seq[5].weight, None, 4, 2, 1)
        self.c8 = SplitConv(seq[8].weight, None, 4, 1, 1)
        self.c11 = SplitConv(seq[11].weight, seq[11].bias, 4, 1, 1, use_bias=True)

    def forward(self, x):
        """x: fp32 NCHW CUDA [B, C <= 16, H, W] -> logits ActBuf (fp32 NHWC, channel 0 valid)"""
        B, C, H, W = x.shape
        din = ActBuf(B, H, W, 3 * split_pad(C), zero=False)
        ops.pack_nchw_split(x.contiguous(), din)
        a, _ = _norm(self.c0(din, act=ACT_LEAKY), ACT_NONE, norm=False)
        a, _ = _norm(self.c2(a), ACT_LEAKY)
        a, _ = _norm(self.c5(a), ACT_LEAKY)
        a, _ = _norm(self.c8(a), ACT_LEAKY)
        return self.c11(a)


def paired_forward_losses(generator, discriminator, input_stack, output_image, l1_weight=100.0):
    """Generator output and the four losses of one train_paired iteration at the CURRENT weights (model.py:615-644;
    no update in between, so the generator's adversarial term uses the same discriminator), all in the parity mode.
    Returns (synthetic fp32 [B,3,H,W], {loss name: float})."""
    G, D = SplitAttentionGenerator(generator), SplitPatchGAN(discriminator)
    synthetic, _ = G.forward(input_stack)
    loss = torch.zeros(4, dtype=torch.float32, device=input_stack.device)
    logits_syn = D.forward(torch.cat((input_stack, synthetic), 1))
    logits_real = D.forward(torch.cat((input_stack, output_image), 1))
    ops.mse_const_loss(logits_real, 1.0, 1.0, 1.0, loss[0:1])
    ops.mse_const_loss(logits_syn, 0.0, 1.0, 1.0, loss[1:2])
    ops.mse_const_loss(logits_syn, 1.0, 1.0, 1.0, loss[2:3])
    ops.l1_loss(synthetic, output_image.contiguous(), l1_weight, 1.0, loss[3:4])
    keys = ("losses_discriminator_real", "losses_discriminator_synthetic", "losses_generator_synthetic",
            "l1_losses_generator_synthetic")
    return synthetic, dict(zip(keys, loss.tolist())), logits_syn
