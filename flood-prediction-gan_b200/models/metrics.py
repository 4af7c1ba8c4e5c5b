"""Image-quality metrics used by Model.calculate_metrics (reference models/model.py:367-371, 404-406), with the call
pattern of the torchmetrics classes the reference instantiates: `metric(preds, target)` returns the batch value as a
0-d tensor, `.reset()` and `.to(device)` exist. Values follow torchmetrics 1.2.0 (requirements.txt:7) for the
configuration the reference uses, data_range=(0, 1):

  PeakSignalNoiseRatio                        10 log10(1 / mean((clamp(p) - clamp(t))^2)), mean over the whole batch
  StructuralSimilarityIndexMeasure            gaussian 11x11 (sigma 1.5), k1 0.01, k2 0.03, 5-pixel border cropped,
                                              mean over channels and positions per image, then over the batch
  MultiScaleStructuralSimilarityIndexMeasure  5 scales (2x2 average pooling between them), betas (0.0448, 0.2856,
                                              0.3001, 0.2363, 0.1333), normalize="relu": per image
                                              prod_k relu(v_k)^beta_k, v_k = contrast sensitivity of scale k, the last
                                              scale's SSIM for k = 4

torchmetrics is an un-vendored dependency that is not installed here: these are restated from its published algorithm and
their parity is UNPINNED (oracle/metrics_oracle.py). LPIPS needs pretrained AlexNet weights and is not provided.
All arithmetic runs in the kernels of csrc/metrics.cu; only 5 numbers per image are combined on the host side.
"""
import torch

from fpgan import ops

MS_SSIM_BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim_size_ok(h, w, n_betas=5, kernel_size=11):
    """the library's precondition: side // (n_betas - 1)^2 must exceed kernel_size - 1 (and side >= 2^n_betas)"""
    div = max(1, n_betas - 1) ** 2
    return min(h, w) >= 2 ** n_betas and min(h, w) // div > kernel_size - 1


class _Metric:
    def to(self, device):
        return self

    def reset(self):
        return None

    def __call__(self, preds, target):
        return self.forward(preds.detach().float(), target.detach().float())


def _clamped(preds, target, data_range):
    lo, hi = data_range
    return preds.clamp(lo, hi), target.clamp(lo, hi), float(hi - lo)


class PeakSignalNoiseRatio(_Metric):
    def __init__(self, data_range=(0, 1)):
        self.data_range = data_range

    def forward(self, preds, target):
        lo, hi = self.data_range
        sse = ops.sq_err_sum(preds, target, clamp=(float(lo), float(hi)))
        mse = sse / preds.numel()
        dr = torch.tensor(float(hi - lo), dtype=torch.float64, device=preds.device)
        return (10.0 * (2 * torch.log10(dr) - torch.log10(mse))).float().squeeze()


class StructuralSimilarityIndexMeasure(_Metric):
    def __init__(self, data_range=(0, 1), sigma=1.5, k1=0.01, k2=0.03):
        self.data_range, self.sigma, self.k1, self.k2 = data_range, sigma, k1, k2

    def forward(self, preds, target):
        p, t, dr = _clamped(preds, target, self.data_range)
        return ops.ssim_stats(p, t, dr, self.k1, self.k2, self.sigma)[:, 0].mean()


class MultiScaleStructuralSimilarityIndexMeasure(_Metric):
    def __init__(self, data_range=(0, 1), sigma=1.5, k1=0.01, k2=0.03, betas=MS_SSIM_BETAS):
        self.data_range, self.sigma, self.k1, self.k2, self.betas = data_range, sigma, k1, k2, betas

    def forward(self, preds, target):
        p, t = preds, target
        if not ms_ssim_size_ok(p.shape[-2], p.shape[-1], len(self.betas)):
            raise ValueError("MS-SSIM: for 5 betas and kernel size 11 the image sides must be larger than 160 "
                             "(side // 16 > 10)")
        per_scale = []
        for k in range(len(self.betas)):
            # the library clamps inside every SSIM evaluation and pools the unclamped images between scales
            pc, tc, dr = _clamped(p, t, self.data_range)
            stats = ops.ssim_stats(pc, tc, dr, self.k1, self.k2, self.sigma)
            last = k == len(self.betas) - 1
            per_scale.append(torch.relu(stats[:, 0 if last else 1]))
            if not last:
                p, t = ops.avgpool2_f32(p), ops.avgpool2_f32(t)
        stack = torch.stack(per_scale)  # [scales, B]
        betas = torch.tensor(self.betas, dtype=stack.dtype, device=stack.device).view(-1, 1)
        return torch.prod(stack ** betas, dim=0).mean()
