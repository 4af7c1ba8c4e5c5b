"""Timing of the image-quality metric kernels on one evaluation batch (B=64, 3x256x256 fp32, BASELINE configs[4] size):
CUDA events, inputs rotated over more than L2. Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
from models import metrics  # noqa: E402


def main():
    g = torch.Generator().manual_seed(0)
    sets = [(torch.rand(64, 3, 256, 256, generator=g).cuda(), torch.rand(64, 3, 256, 256, generator=g).cuda())
            for _ in range(4)]  # 4 x 100 MB
    fns = {"PSNR": metrics.PeakSignalNoiseRatio(), "SSIM": metrics.StructuralSimilarityIndexMeasure(),
           "MS-SSIM": metrics.MultiScaleStructuralSimilarityIndexMeasure()}
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    out = {"batch": 64, "image": "3x256x256 fp32", "hbm_peak_GBps": hbm}
    nbytes = 2 * 64 * 3 * 256 * 256 * 4
    for name, m in fns.items():
        for p, t in sets:
            m(p, t)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for p, t in sets:
                m(p, t)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / len(sets))
        ts.sort()
        ms = ts[len(ts) // 2]
        out[name] = {"us_per_batch": ms * 1e3, "GBps_algorithmic": nbytes / ms / 1e6,
                     "frac_of_hbm": nbytes / ms / 1e6 / hbm}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
