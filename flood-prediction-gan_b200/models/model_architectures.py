"""Drop-in generator / discriminator classes with the reference's names, constructor signatures, sub-module names
and state_dict layout (reference: models/model_architectures.py), computing through the sm_100a kernel library.

The torch.nn layers instantiated here are PARAMETER CONTAINERS only (same construction order and RNG consumption as
the reference, so `Model(seed=...)` initialises identical weights and reference checkpoints load unchanged); their
own forward() is never called. forward() hands the whole network to a native executor (fpgan.networks) wrapped in
one torch.autograd.Function, so `loss.backward()` / optimisers / requires_grad toggling work as in the reference.
There is no CPU or eager-PyTorch fallback: inputs must live on a CUDA device.
"""
import torch
from torch import nn

from fpgan import networks, ops


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: input is on {x.device}; this implementation only runs on a CUDA (sm_100a) device "
                           "and has no CPU fallback")


class _NetFunction(torch.autograd.Function):
    """One autograd node per network call; parameters are inputs so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, owner, x, *params):
        net = owner._executor()
        x32 = x.detach().float().contiguous()
        out, tape = net.forward(x32)
        ctx.net, ctx.tape, ctx.owner = net, tape, owner
        ctx.n_params = len(params)
        ctx.in_channels = x.shape[1]
        owner._after_forward(tape)
        return out

    @staticmethod
    def backward(ctx, dout):
        net = ctx.net
        need_dx = ctx.needs_input_grad[1]
        need_dw = any(ctx.needs_input_grad[2:])
        named = net.named_params()
        grads = networks.Grads(named) if need_dw else None
        dx = ctx.owner._run_backward(net, ctx.tape, dout.contiguous().float(), grads, need_dx)
        ctx.tape = None
        if need_dw:
            gl = tuple(grads[n] if ctx.needs_input_grad[2 + i] else None for i, (n, _) in enumerate(named))
        else:
            gl = (None,) * ctx.n_params
        return (None, dx) + gl


class _NativeModule(nn.Module):
    _executor_cls = None

    def _executor(self):
        p = next(self.parameters())
        key = (p.device, p.data_ptr())
        if getattr(self, "_exec_key", None) != key:
            if not p.is_cuda:
                raise RuntimeError(f"{type(self).__name__}: parameters are on {p.device}; move the module to a CUDA "
                                   "device (no CPU fallback)")
            object.__setattr__(self, "_exec", type(self)._executor_cls(self))
            object.__setattr__(self, "_exec_key", key)
        return self._exec

    def _after_forward(self, tape):
        pass

    def forward(self, x):
        _require_cuda(x, type(self).__name__)
        params = [p for _, p in self.named_parameters()]
        return _NetFunction.apply(self, x, *params)


# ------------------------------------------------------------------------------------------------ generators
class _AttentionGeneratorBase(_NativeModule):
    _executor_cls = networks.AttentionGeneratorNet

    def __init__(self, input_channels, block_cls):
        super().__init__()
        self.input_channels = input_channels
        self.last_attention_mask = None
        self.conv1 = nn.Conv2d(input_channels, 64, kernel_size=7, stride=1, padding=0)
        self.conv1_norm = nn.InstanceNorm2d(64)
        self.conv2 = nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1)
        self.conv2_norm = nn.InstanceNorm2d(128)
        self.conv3 = nn.Conv2d(128, 256, kernel_size=3, stride=2, padding=1)
        self.conv3_norm = nn.InstanceNorm2d(256)
        self.resnet_blocks = nn.Sequential(*[block_cls(channel=256, kernel=3, stride=1, padding=1) for _ in range(9)])
        for branch, out_ch, k in (("content", 27, 7), ("attention", 10, 1)):
            setattr(self, f"deconv1_{branch}", nn.ConvTranspose2d(256, 128, 3, 2, 1, 1))
            setattr(self, f"deconv1_norm_{branch}", nn.InstanceNorm2d(128))
            setattr(self, f"deconv2_{branch}", nn.ConvTranspose2d(128, 64, 3, 2, 1, 1))
            setattr(self, f"deconv2_norm_{branch}", nn.InstanceNorm2d(64))
            setattr(self, f"deconv3_{branch}", nn.Conv2d(64, out_ch, k, 1, 0))
        self.tanh = nn.Tanh()
        self.softmax = nn.Softmax(dim=1)

    def _after_forward(self, tape):
        self.last_attention_mask = tape["mask"]  # background attention a_10, [B, H, W] (reference :396)

    def _run_backward(self, net, tape, dout, grads, need_dx):
        return net.backward(tape, grads, dout_nchw=dout, need_dx=need_dx)


class _ResidualBlockParams(nn.Module):
    """Parameter container of one residual block (reference :402-418 / :260-276); executed by the generator."""

    def __init__(self, channel, kernel, stride, padding):
        super().__init__()
        self.padding = padding
        self.conv1 = nn.Conv2d(channel, channel, kernel, stride, 0)
        self.conv1_norm = nn.InstanceNorm2d(channel)
        self.conv2 = nn.Conv2d(channel, channel, kernel, stride, 0)
        self.conv2_norm = nn.InstanceNorm2d(channel)

    def forward(self, x):
        raise RuntimeError("residual blocks are executed by the enclosing generator's native executor")


class PairedAttentionBlock(_ResidualBlockParams):
    pass


class AttentionGANBlock(_ResidualBlockParams):
    pass


class PairedAttentionGenerator(_AttentionGeneratorBase):
    def __init__(self, input_channels):
        super().__init__(input_channels, PairedAttentionBlock)


class AttentionGANGenerator(_AttentionGeneratorBase):
    def __init__(self, input_channels):
        super().__init__(input_channels, AttentionGANBlock)


class CycleGANBlock(nn.Module):
    """Parameter container of one CycleGAN residual block (reference :122-134); executed by the generator."""

    def __init__(self, dim):
        super().__init__()
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=True),
                                        nn.InstanceNorm2d(dim), nn.ReLU(True), nn.ReflectionPad2d(1),
                                        nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=True),
                                        nn.InstanceNorm2d(dim))

    def forward(self, x):
        raise RuntimeError("residual blocks are executed by the enclosing generator's native executor")


class CycleGANGenerator(_NativeModule):
    """reference :91-120; `model` keeps the reference's Sequential indices (state_dict keys model.1, model.4, ...)."""
    _executor_cls = networks.CycleGANGeneratorNet

    def __init__(self, input_channels):
        super().__init__()
        seq = [nn.ReflectionPad2d(3), nn.Conv2d(input_channels, 64, kernel_size=7, padding=0, bias=True),
               nn.InstanceNorm2d(64), nn.ReLU(True)]
        for mult in (1, 2):
            seq += [nn.Conv2d(64 * mult, 128 * mult, kernel_size=3, stride=2, padding=1, bias=True),
                    nn.InstanceNorm2d(128 * mult), nn.ReLU(True)]
        seq += [CycleGANBlock(dim=256) for _ in range(9)]
        for mult in (4, 2):
            seq += [nn.ConvTranspose2d(64 * mult, 32 * mult, kernel_size=3, stride=2, padding=1, output_padding=1,
                                       bias=True), nn.InstanceNorm2d(32 * mult), nn.ReLU(True)]
        seq += [nn.ReflectionPad2d(3), nn.Conv2d(64, 3, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*seq)

    def _run_backward(self, net, tape, dout, grads, need_dx):
        return net.backward(tape, grads, dout, need_dx=need_dx)


# ------------------------------------------------------------------------------------------------ discriminators
class _InstanceNormPatchGAN(_NativeModule):
    _executor_cls = networks.PatchGANNet

    def __init__(self, in_channels):
        super().__init__()
        seq = [nn.Conv2d(in_channels, 64, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        prev = 64
        for mult in (2, 4):
            seq += [nn.Conv2d(prev, 64 * mult, kernel_size=4, stride=2, padding=1, bias=True),
                    nn.InstanceNorm2d(64 * mult), nn.LeakyReLU(0.2, True)]
            prev = 64 * mult
        seq += [nn.Conv2d(prev, 512, kernel_size=4, stride=1, padding=1, bias=True), nn.InstanceNorm2d(512),
                nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(512, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*seq)

    def _run_backward(self, net, tape, dout, grads, need_dx):
        dl = ops.ActBuf(dout.shape[0], dout.shape[2], dout.shape[3], 16, zero=False)
        ops.pack_nchw(dout.contiguous(), dl, 0, zero_rest=True)  # fp32 NCHW logit gradient -> bf16 NHWC, native kernel
        dd = net.backward(tape, dl, grads, need_dx)
        if not need_dx:
            return None
        c = tape["c_in"]
        dx = torch.empty(dd.n, c, dd.h, dd.w, dtype=torch.float32, device=dout.device)
        ops.unpack_nchw(dd, dx, 0)
        return dx

    def _after_forward(self, tape):
        tape["c_in"] = self.model[0].weight.shape[1]


class PairedAttentionDiscriminator(_InstanceNormPatchGAN):
    def __init__(self, input_channels):
        super().__init__(input_channels + 3)


class AttentionGANDiscriminator(_InstanceNormPatchGAN):
    def __init__(self, input_channels):
        super().__init__(input_channels)


class CycleGANDiscriminator(_InstanceNormPatchGAN):
    def __init__(self, input_channels):
        super().__init__(input_channels)


# ------------------------------------------------------------------------------------------------ Pix2Pix
class Pix2PixBlock(nn.Module):
    """Parameter container of one U-Net block with the reference's Sequential layout (reference :25-63), so that the
    nested state_dict keys (model.model.1.model.3...) and the construction / initialisation order are identical."""

    def __init__(self, outer_nc, inner_nc, input_nc, submodule, outermost, innermost, use_dropout):
        super().__init__()
        self.outermost = outermost
        if input_nc is None:
            input_nc = outer_nc
        downconv = nn.Conv2d(input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=False)
        downrelu, uprelu = nn.LeakyReLU(0.2, True), nn.ReLU(True)
        downnorm, upnorm = nn.BatchNorm2d(inner_nc), nn.BatchNorm2d(outer_nc)
        if outermost:
            upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downconv, submodule, uprelu, upconv, nn.Tanh()]
        elif innermost:
            upconv = nn.ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=False)
            model = [downrelu, downconv, uprelu, upconv, upnorm]
        else:
            upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=False)
            model = [downrelu, downconv, downnorm, submodule, uprelu, upconv, upnorm]
            if use_dropout:
                model.append(nn.Dropout(0.5))
        self.model = nn.Sequential(*model)

    def forward(self, x):
        raise RuntimeError("U-Net blocks are executed by the enclosing generator's native executor")


class Pix2PixGenerator(_NativeModule):
    """reference :9-23. Always computes in training mode (batch statistics, dropout), as the reference does -- it
    never calls .eval()."""
    _executor_cls = networks.Pix2PixGeneratorNet

    def __init__(self, input_channels):
        super().__init__()
        block = Pix2PixBlock(512, 512, None, None, False, True, False)
        for _ in range(3):
            block = Pix2PixBlock(512, 512, None, block, False, False, True)
        block = Pix2PixBlock(256, 512, None, block, False, False, False)
        block = Pix2PixBlock(128, 256, None, block, False, False, False)
        block = Pix2PixBlock(64, 128, None, block, False, False, False)
        self.model = Pix2PixBlock(3, 64, input_channels, block, True, False, False)

    def _run_backward(self, net, tape, dout, grads, need_dx):
        return net.backward(tape, grads, dout, need_dx=need_dx)


class Pix2PixDiscriminator(_NativeModule):
    """reference :65-85: BatchNorm PatchGAN on cat(input stack, image)."""
    _executor_cls = networks.PatchGANBatchNormNet

    def __init__(self, input_channels):
        super().__init__()
        seq = [nn.Conv2d(input_channels + 3, 64, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        prev = 64
        for mult in (2, 4):
            seq += [nn.Conv2d(prev, 64 * mult, kernel_size=4, stride=2, padding=1, bias=False),
                    nn.BatchNorm2d(64 * mult), nn.LeakyReLU(0.2, True)]
            prev = 64 * mult
        seq += [nn.Conv2d(prev, 512, kernel_size=4, stride=1, padding=1, bias=False), nn.BatchNorm2d(512),
                nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(512, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*seq)

    def _run_backward(self, net, tape, dout, grads, need_dx):
        dl = ops.ActBuf(dout.shape[0], dout.shape[2], dout.shape[3], 16, zero=False)
        ops.pack_nchw(dout.contiguous(), dl, 0, zero_rest=True)  # fp32 NCHW logit gradient -> bf16 NHWC, native kernel
        dd = net.backward(tape, dl, grads, need_dx)
        if not need_dx:
            return None
        c = self.model[0].weight.shape[1]
        dx = torch.empty(dd.n, c, dd.h, dd.w, dtype=torch.float32, device=dout.device)
        ops.unpack_nchw(dd, dx, 0)
        return dx


# ------------------------------------------------------------------------------------------------ segmentation U-Net
class DoubleConv(nn.Module):
    """parameter container (reference :541-552)"""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid_channels = mid_channels or out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class Up(nn.Module):
    def __init__(self, in_channels, out_channels, bilinear=False):
        super().__init__()
        if bilinear:
            raise NotImplementedError("the reference instantiates UNet(bilinear=False) only")
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels)


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)


class UNet(nn.Module):
    """Segmentation U-Net of calculate_metrics (reference :508-586), inference through the native executor. Like the
    reference it always runs BatchNorm with batch statistics (no .eval() anywhere in the reference)."""

    def __init__(self, n_channels=3, n_classes=1, bilinear=False):
        super().__init__()
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1, self.down2 = Down(64, 128), Down(128, 256)
        self.down3, self.down4 = Down(256, 512), Down(512, 1024)
        self.up1, self.up2 = Up(1024, 512), Up(512, 256)
        self.up3, self.up4 = Up(256, 128), Up(128, 64)
        self.outc = OutConv(64, n_classes)

    def _executor(self):
        p = next(self.parameters())
        key = (p.device, p.data_ptr())
        if getattr(self, "_exec_key", None) != key:
            if not p.is_cuda:
                raise RuntimeError("UNet: parameters must be on a CUDA device (no CPU fallback)")
            object.__setattr__(self, "_exec", networks.UNetNet(self))
            object.__setattr__(self, "_exec_key", key)
        return self._exec

    @torch.no_grad()
    def forward(self, x):
        _require_cuda(x, "UNet")
        out, _ = self._executor().forward(x.detach().float().contiguous())
        return out


def flood_masks_and_counts(seg_model, generated, ground_truth):
    """model.py:397-418: rescale both images to [0, 1], segment, threshold with the bit-exact (sigmoid > 0.5) kernel,
    count TP / FP / TN / FN on the device. Returns (output_mask, true_mask, counts int64[4])."""
    gt = torch.clamp((ground_truth + 1) * 0.5, min=0, max=1)
    gen = torch.clamp((generated + 1) * 0.5, min=0, max=1)
    masks = []
    for img in (gen, gt):
        logits = seg_model(img).contiguous()
        m = torch.empty_like(logits)
        ops.flood_mask(logits, m)
        masks.append(m)
    counts = torch.zeros(4, dtype=torch.int64, device=generated.device)
    ops.confusion_counts(masks[0].reshape(-1), masks[1].reshape(-1), counts)
    return masks[0], masks[1], counts
