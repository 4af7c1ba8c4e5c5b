"""CPU oracle of the image-quality metrics of calculate_metrics (reference models/model.py:367-371, 404-406).
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic lives in torchmetrics==1.2.0 (requirements.txt:7), an un-vendored dependency that is
not installed in this environment, and the reference holds no golden values for it. This file restates the published
algorithm of torchmetrics.functional.image (psnr.py, ssim.py) for the configuration the reference uses
(data_range=(0, 1), defaults otherwise), with plain torch CPU ops in the library's own formulation -- reflect padding,
one 2-D gaussian depthwise convolution over [p, t, p*p, t*t, p*t], cropping, per-image means -- so that the device
kernels (which use a separable window on the cropped region only) are checked against an independent evaluation.
"""
import torch
import torch.nn.functional as F

BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def psnr(preds, target, data_range=(0.0, 1.0)):
    p, t = preds.clamp(*data_range).double(), target.clamp(*data_range).double()
    mse = ((p - t) ** 2).sum() / p.numel()
    dr = torch.tensor(data_range[1] - data_range[0], dtype=torch.float64)
    return float(10.0 * (2 * torch.log10(dr) - torch.log10(mse)))


def _gaussian_2d(channels, size=11, sigma=1.5, dtype=torch.float32):
    dist = torch.arange((1 - size) / 2, (1 + size) / 2, 1, dtype=dtype)
    g = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    g = (g / g.sum()).unsqueeze(0)
    return torch.matmul(g.t(), g).expand(channels, 1, size, size)


def ssim_and_cs(preds, target, data_range=(0.0, 1.0), k1=0.01, k2=0.03, size=11, sigma=1.5):
    """per-image (mean SSIM, mean contrast sensitivity): torchmetrics _ssim_update(return_contrast_sensitivity=True)"""
    p, t = preds.clamp(*data_range).double(), target.clamp(*data_range).double()
    dr = data_range[1] - data_range[0]
    c1, c2 = (k1 * dr) ** 2, (k2 * dr) ** 2
    b, c = p.shape[:2]
    pad = (size - 1) // 2
    p, t = F.pad(p, (pad,) * 4, mode="reflect"), F.pad(t, (pad,) * 4, mode="reflect")
    kernel = _gaussian_2d(c, size, sigma, dtype=torch.float64)
    out = F.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel, groups=c).split(b)
    mu_pp, mu_tt, mu_pt = out[0] ** 2, out[1] ** 2, out[0] * out[1]
    sig_p, sig_t, sig_pt = out[2] - mu_pp, out[3] - mu_tt, out[4] - mu_pt
    upper, lower = 2 * sig_pt + c2, sig_p + sig_t + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower)
    crop = (..., slice(pad, -pad), slice(pad, -pad))
    return full[crop].reshape(b, -1).mean(-1), (upper / lower)[crop].reshape(b, -1).mean(-1)


def ssim(preds, target, **kw):
    return float(ssim_and_cs(preds, target, **kw)[0].mean())


def ms_ssim(preds, target, betas=BETAS, **kw):
    p, t = preds, target
    vals = []
    for k in range(len(betas)):
        s, cs = ssim_and_cs(p, t, **kw)
        vals.append(torch.relu(s if k == len(betas) - 1 else cs))
        p, t = F.avg_pool2d(p, (2, 2)), F.avg_pool2d(t, (2, 2))
    stack = torch.stack(vals)
    w = torch.tensor(betas, dtype=stack.dtype).view(-1, 1)
    return float(torch.prod(stack ** w, dim=0).mean())
