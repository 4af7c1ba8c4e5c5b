"""Input pipeline of the training / evaluation loops (reference: models/data.py:11-150, models/utils.py:19-67).

Same interface as the reference -- `create_flood_dataset(...)` returns (train, validation, test) loaders that yield
(input_stack [B,C,h,w], output_image [B,3,h,w], names) -- re-designed for a 180 GB GPU:

  reference (per sample, per epoch, one CPU thread)      here
  ------------------------------------------------       -------------------------------------------------------
  tifffile.imread of the 1024x1024x9 fp32 stack          decoded ONCE per (file, version), uploaded, resized by
  torchvision bicubic anti-aliased Resize                `fpg_resize_bicubic_aa` (channel selection and the left-
  (repeated for every one of the `crop` windows)         right flip fused) and kept RESIDENT in HBM as fp32 CHW
  crop window, Normalize(0.5, 0.5)                       `fpg_tile_gather`: one launch per batch and tensor
  DataLoader(shuffle=True, pin_memory=True)              the same RandomSampler permutation (global torch RNG)

The whole resized dataset (2336 pairs x 12.6 MB at resize=512) is 29 GB. Batches are produced on the device: the
training step starts from HBM-resident inputs, there is no host->device copy per step.

TIFF decoding itself is storage format, outside the accelerated path: `decoder(path) -> HWC float32 array` defaults to
tifffile.imread when that package is installed; any callable can be injected (the tests inject synthetic stacks).
"""
import math
import os

import numpy as np
import torch

TOPOGRAPHY_CHANNELS = {"all": 9, "map": 6, "dem": 4, "flow": 4, "river": 4, None: 3}
# channels of the decoded 9-channel stack kept per topography (models/utils.py:30-39)
TOPOGRAPHY_CHANNEL_MAP = {"all": tuple(range(9)), "dem": (0, 1, 2, 3), "flow": (0, 1, 2, 4), "river": (0, 1, 2, 5),
                          "map": (0, 1, 2, 6, 7, 8), None: (0, 1, 2)}
METADATA_CSV = os.path.join("metadata", "dataset_split.csv")  # relative to the cwd, as in the reference (data.py:90)


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


# ---------------------------------------------------------------------------------------------- dataset split
def _relabel(frame, mask, split):
    frame.loc[mask, "split"] = split


def _leave_one_disaster_out(pd, table, pool, train, validation, flip_source):
    """`harveyflorence` / `harveyonflorence` (data.py:99-118): train on `train` disasters (their test images are added
    once more as a flipped version), validate AND test on the `validation` disaster, never on flipped images."""
    frame = table[pool(table)].copy()
    extra = frame[flip_source(frame) & (frame["split"] == "test")].copy()
    extra["version"] = "flipped"
    frame = pd.concat([frame, extra], axis=0)
    _relabel(frame, train(frame), "train")
    _relabel(frame, validation(frame), "validation")
    as_test = frame[validation(frame)].copy()
    as_test["split"] = "test"
    frame = pd.concat([frame, as_test], axis=0).reset_index(drop=True)
    held_out = (frame["split"] == "test") | (frame["split"] == "validation")
    return frame.drop(frame[held_out & (frame["version"] == "flipped")].index)


def determine_flood_dataset(subset, dem, crop=None, metadata_csv=None):
    """The image files (file name, version[, crop index]) of each split -- reference data.py:84-146, same selection,
    same pandas shuffles (random_state 47), hence the same order."""
    import pandas as pd
    table = pd.read_csv(metadata_csv or METADATA_CSV)
    key = subset.lower()
    harvey = lambda f: f["disaster"] == "hurricane-harvey"  # noqa: E731
    florence = lambda f: f["disaster"] == "hurricane-florence"  # noqa: E731
    if key in ("usa", "india"):
        frame = table[table["country"] == key].copy()
    elif key in ("hurricane-harvey", "hurricane-florence", "midwest-flooding", "nepal-flooding"):
        frame = table[table["disaster"] == key].copy()
    elif key == "harveyflorence":
        frame = _leave_one_disaster_out(pd, table, pool=lambda f: f["country"] == "usa",
                                        train=lambda f: harvey(f) | florence(f),
                                        validation=lambda f: f["disaster"] == "midwest-flooding",
                                        flip_source=lambda f: harvey(f) | florence(f))
    elif key == "harveyonflorence":
        frame = _leave_one_disaster_out(pd, table, pool=lambda f: harvey(f) | florence(f), train=harvey,
                                        validation=florence, flip_source=harvey)
    elif key == "testing":
        frame = table[harvey(table)].copy()
        frame = frame[frame["version"] == "original"].sample(n=50, random_state=47)
    elif key == "all":
        frame = table.copy()
    else:
        raise NotImplementedError("Unrecognised dataset subset name")
    if dem not in ("best", "same"):
        raise NotImplementedError("Unrecognised DEM name - provide 'best' or 'same'")
    frame["file_name"] = frame["image"] + "_" + frame[f"{dem}_DEM"] + ".tif"
    frame = frame.sample(frac=1, random_state=47)
    columns = ["file_name", "version"]
    if crop:
        frame = pd.concat([frame.assign(crop=i) for i in range(crop)])
        columns.append("crop")
    out = {}
    for name in ("train", "validation", "test"):
        part = frame[frame["split"] == name]
        out[name] = list(zip(*(part[c] for c in columns)))
    return out


# ---------------------------------------------------------------------------------------------- resident images
def _default_decoder(path):
    try:
        import tifffile
    except ImportError as e:  # storage format: not part of the accelerated path
        raise ImportError("reading the dataset needs the `tifffile` package (or pass decoder=callable returning the "
                          "HWC float32 array of a path)") from e
    return tifffile.imread(path)


def resize_output_size(h, w, size):
    """torchvision Resize(int): the smaller edge becomes `size`, the other keeps the aspect ratio"""
    if not size:
        return h, w
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


class ResidentImages:
    """(file, version) -> resized fp32 CHW input / output images in HBM, filled on first use."""

    def __init__(self, path, topography, resize, decoder=None, device="cuda"):
        self.path, self.topography, self.resize = path, topography, resize
        self.decoder = decoder or _default_decoder
        self.device = device
        self.images = {}

    def get(self, file_name, version):
        key = (file_name, version)
        hit = self.images.get(key)
        if hit is None:
            hit = self.images[key] = self._load(file_name, version == "flipped")
        return hit

    def _load(self, file_name, flipped):
        from fpgan import ops
        stem = file_name[:-8]  # "<image>_<dem>.tif" -> "<image>" (data.py:61)
        pair = []
        for folder, name, cmap in (("dataset_input", file_name, TOPOGRAPHY_CHANNEL_MAP[self.topography]),
                                   ("dataset_output", stem + ".tif", (0, 1, 2))):
            host = np.ascontiguousarray(self.decoder(f"{self.path}/{folder}/{name}"), dtype=np.float32)
            dev = torch.from_numpy(host).to(self.device, non_blocking=False)
            oh, ow = resize_output_size(host.shape[0], host.shape[1], self.resize)
            pair.append(ops.resize_bicubic_aa(dev, cmap, oh, ow, flip_w=flipped))
        return tuple(pair)

    def bytes(self):
        return sum(a.numel() * 4 + b.numel() * 4 for a, b in self.images.values())


class FloodDataset:
    """Reference FloodDataset (data.py:46-82) over resident images: item = (input [C,h,w], output [3,h,w], name),
    device tensors."""

    def __init__(self, dataset_subset, dataset_dem, split, path, topography, resize, crop, decoder=None,
                 metadata_csv=None, device="cuda"):
        self.data_files = determine_flood_dataset(dataset_subset, dataset_dem, crop, metadata_csv)[split]
        self.resize, self.path, self.crop, self.topography = resize, path, crop, topography
        self.store = ResidentImages(path, topography, resize, decoder, device)
        self.device = device

    def __len__(self):
        return len(self.data_files)

    def item_name(self, index):
        entry = self.data_files[index]
        name = entry[0][:-8]
        return f"{name}_{entry[2]}" if self.crop else name

    def gather(self, indices):
        """batch of items: one gather launch per tensor"""
        from fpgan import ops
        entries = [self.data_files[i] for i in indices]
        pairs = [self.store.get(e[0], e[1]) for e in entries]
        crops = [int(e[2]) if self.crop else 0 for e in entries]
        div = int(math.sqrt(self.crop)) if self.crop else 1
        table = torch.tensor([[p[0].data_ptr() for p in pairs], [p[1].data_ptr() for p in pairs]], dtype=torch.int64)
        table = table.to(self.device, non_blocking=True)
        crop_t = torch.tensor(crops, dtype=torch.int32).to(self.device, non_blocking=True)
        outs = []
        for row, ref in ((table[0], pairs[0][0]), (table[1], pairs[0][1])):
            c, h, w = ref.shape
            out = torch.empty(len(indices), c, h // div, w // div, dtype=torch.float32, device=self.device)
            outs.append(ops.tile_gather(row, crop_t, c, h, w, div, out))
        return outs[0], outs[1], [self.item_name(i) for i in indices]

    def __getitem__(self, index):
        x, y, names = self.gather([index])
        return x[0], y[0], names[0]


class DeviceLoader:
    """DataLoader(dataset, batch_size, shuffle=True) semantics (data.py:28-43) without worker processes or host copies:
    a new permutation per epoch drawn exactly like torch's RandomSampler (a seed taken from the global torch RNG, after
    the one DataLoader's iterator takes for its workers), last batch kept. With world_size > 1 every rank draws the same
    permutation and takes rows [rank*B, (rank+1)*B) of each global batch of world_size*B samples; a ragged tail is
    completed by wrapping around to the start of the permutation (torch's DistributedSampler does the same), so that
    every rank takes the same number of equally sized batches -- the fused step issues NCCL all-reduces (captured in
    a CUDA graph keyed by the batch shape) that every rank must join, with equal weight 1/world_size."""

    def __init__(self, dataset, batch_size=1, shuffle=True, rank=0, world_size=1):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle
        self.rank, self.world_size = rank, world_size

    def __len__(self):
        return (len(self.dataset) + self.batch_size * self.world_size - 1) // (self.batch_size * self.world_size)

    def order(self):
        n = len(self.dataset)
        torch.empty((), dtype=torch.int64).random_()  # DataLoader iterator: base seed of the (absent) workers
        if not self.shuffle:
            return list(range(n))
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g).tolist()

    def rank_batches(self, order):
        """index lists of this rank's batches for one epoch"""
        gb = self.batch_size * self.world_size
        if self.world_size > 1 and order and len(order) % gb:
            pad = gb - len(order) % gb
            order = order + (order * (pad // len(order) + 1))[:pad]
        lo, hi = self.rank * self.batch_size, (self.rank + 1) * self.batch_size
        return [order[start:start + gb][lo:hi] for start in range(0, len(order), gb)]

    def __iter__(self):
        for chunk in self.rank_batches(self.order()):
            yield self.dataset.gather(chunk)


def create_flood_dataset(dataset_subset="all", dataset_dem="best", data_path=None, topography="all", resize=None,
                         crop=None, batch_size=1, num_workers=0, decoder=None, metadata_csv=None, rank=0, world_size=1):
    """Returns (train, validation, test) loaders (reference data.py:11-44). Without a dataset on disk (no
    <data_path>/dataset_input and no injected decoder) empty loaders are returned and the caller assigns its own
    iterable (Model.train_loader = ...), as the benchmarks and parity tests do."""
    have_files = data_path is not None and os.path.isdir(os.path.join(str(data_path), "dataset_input"))
    if decoder is None and not have_files:
        return [], [], []
    loaders = []
    for split in ("train", "validation", "test"):
        ds = FloodDataset(dataset_subset, dataset_dem, split, data_path, topography, resize, crop, decoder,
                          metadata_csv)
        loaders.append(DeviceLoader(ds, batch_size, shuffle=True, rank=rank, world_size=world_size))
    return tuple(loaders)
