"""Input contract of the training step (reference: models/data.py:11-44, models/utils.py:19-67).

The reference streams 9-channel float32 TIFF stacks from disk (tifffile) one tile at a time. Dataset construction
and decoding are outside the hot path this repository accelerates (SURVEY.md section 8f-2 ranks it "next"): this
module only fixes the loader contract -- an iterable of (input_stack [B,C,H,W], output_image [B,3,H,W], names)
-- and provides the synthetic loader used by the benchmarks and parity tests.
"""
import os

import torch

TOPOGRAPHY_CHANNELS = {"all": 9, "map": 6, "dem": 4, "flow": 4, "river": 4, None: 3}


class SyntheticLoader:
    """i.i.d. uniform [-1, 1] tiles, generator seeded with 1000 + step; rank r of W takes rows
    [r*B, (r+1)*B) of the global batch so that a sharded run sees the same global batch as a single process."""

    def __init__(self, steps, batch, channels=9, size=256, rank=0, world_size=1, pin=True):
        self.steps, self.batch, self.channels, self.size = steps, batch, channels, size
        self.rank, self.world_size, self.pin = rank, world_size, pin

    def __len__(self):
        return self.steps

    def __iter__(self):
        gb = self.batch * self.world_size
        for step in range(self.steps):
            g = torch.Generator().manual_seed(1000 + step)
            x = torch.rand(gb, self.channels, self.size, self.size, generator=g) * 2 - 1
            y = torch.rand(gb, 3, self.size, self.size, generator=g) * 2 - 1
            lo = self.rank * self.batch
            x, y = x[lo:lo + self.batch].contiguous(), y[lo:lo + self.batch].contiguous()
            if self.pin and torch.cuda.is_available():
                x, y = x.pin_memory(), y.pin_memory()
            yield x, y, tuple(f"synthetic_{step}_{i}" for i in range(self.batch))


def create_flood_dataset(dataset_subset="all", dataset_dem="best", data_path=None, topography="all", resize=256,
                         crop=None):
    """Returns (train, val, test) loaders. Without a data_path there is no dataset on disk: empty loaders are
    returned and the caller injects its own iterable (Model.train_loader = ...), as the parity tests do."""
    if data_path is None or not os.path.isdir(os.path.join(str(data_path), "dataset_input")):
        return [], [], []
    raise NotImplementedError(
        "reading the xBD-derived TIFF stacks is outside the accelerated hot path (SURVEY.md section 8f-2); "
        "assign Model.train_loader an iterable of (input_stack, output_image, names) batches")
