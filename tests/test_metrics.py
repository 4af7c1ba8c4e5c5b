"""Image-quality metrics of calculate_metrics (reference models/model.py:367-371, 404-406).
torchmetrics is absent, so the oracle (oracle/metrics_oracle.py) is UNPINNED against the library itself; it is anchored
instead on the published definitions: an independent scipy evaluation of the Wang et al. SSIM with the same gaussian
window, closed-form PSNR values, and the algebraic properties of the indices. The GPU tests compare the kernels
(through the C ABI) with the oracle; tolerance 2e-5 absolute (fp32 window sums against float64)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as MO  # noqa: E402


def pair(seed, b=2, c=3, size=192, noise=0.1):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(b, c, size, size, generator=g)
    return a, (a + noise * torch.randn(b, c, size, size, generator=g)).clamp(0, 1)


def test_oracle_psnr_closed_form():
    a = torch.full((1, 3, 8, 8), 0.25)
    assert abs(MO.psnr(a, a + 0.1) - 20.0) < 1e-5            # mse 0.01 -> 20 dB
    assert abs(MO.psnr(a, a + 0.5) - 10 * np.log10(4.0)) < 1e-5
    assert abs(MO.psnr(a - 1.0, a) - 10 * np.log10(1 / 0.0625)) < 1e-5  # inputs are clamped to the data range first


def test_oracle_ssim_matches_independent_scipy_evaluation():
    """Wang et al. SSIM with a gaussian window (sigma 1.5, truncated at radius 5), population covariances, 5-pixel
    border dropped -- evaluated with scipy.ndimage instead of a depthwise convolution"""
    from scipy.ndimage import gaussian_filter
    a, b = pair(1, b=1, c=1, size=96)
    x, y = a[0, 0].double().numpy(), b[0, 0].double().numpy()
    f = lambda im: gaussian_filter(im, sigma=1.5, truncate=3.5, mode="reflect")  # noqa: E731
    ux, uy = f(x), f(y)
    vx, vy, vxy = f(x * x) - ux * ux, f(y * y) - uy * uy, f(x * y) - ux * uy
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    smap = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    want = smap[5:-5, 5:-5].mean()
    assert abs(MO.ssim(a, b) - want) < 1e-9


def test_oracle_index_properties():
    a, b = pair(2)
    assert abs(MO.ssim(a, a) - 1.0) < 1e-12 and abs(MO.ms_ssim(a, a) - 1.0) < 1e-12
    assert abs(MO.ssim(a, b) - MO.ssim(b, a)) < 1e-12          # symmetric
    s1, s2 = MO.ssim(a, b), MO.ssim(a, pair(2, noise=0.3)[1])
    assert s2 < s1 < 1.0                                         # more distortion, lower index
    assert 0.0 < MO.ms_ssim(a, b) < 1.0
    per_image = MO.ssim_and_cs(a, b)[0]
    assert abs(float(per_image.mean()) - MO.ssim(a, b)) < 1e-12 and per_image.shape == (2,)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 3, 192, 192), (1, 3, 256, 256), (3, 1, 171, 203), (1, 2, 176, 181)])
def test_device_metrics_match_oracle(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from models import metrics
    b, c, h, w = shape
    g = torch.Generator().manual_seed(h)
    a = torch.rand(b, c, h, w, generator=g)
    t = (a + 0.15 * torch.randn(b, c, h, w, generator=g))  # not clamped: the metrics clamp to the data range themselves
    ad, td = a.cuda(), t.cuda()
    assert abs(metrics.PeakSignalNoiseRatio(data_range=(0, 1))(ad, td).item() - MO.psnr(a, t)) < 1e-4
    assert abs(metrics.StructuralSimilarityIndexMeasure(data_range=(0, 1))(ad, td).item() - MO.ssim(a, t)) < 2e-5
    if min(h, w) // 16 > 10:  # the library's size precondition for 5 scales
        got = metrics.MultiScaleStructuralSimilarityIndexMeasure(data_range=(0, 1))(ad, td).item()
        assert abs(got - MO.ms_ssim(a, t)) < 2e-5
    else:
        with pytest.raises(ValueError):
            metrics.MultiScaleStructuralSimilarityIndexMeasure(data_range=(0, 1))(ad, td)


@pytest.mark.gpu
def test_device_metrics_properties_and_pooling():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from fpgan import ops
    from models import metrics
    g = torch.Generator().manual_seed(9)
    a = torch.rand(4, 3, 256, 256, generator=g).cuda()
    assert abs(metrics.StructuralSimilarityIndexMeasure()(a, a).item() - 1.0) < 1e-6
    assert abs(metrics.MultiScaleStructuralSimilarityIndexMeasure()(a, a).item() - 1.0) < 1e-5
    assert torch.equal(ops.avgpool2_f32(a), torch.nn.functional.avg_pool2d(a, (2, 2)))
    stats = ops.ssim_stats(a, a.flip(0))
    assert stats.shape == (4, 2) and torch.allclose(stats, stats.flip(0), atol=1e-6)  # symmetric per image pair
