"""Input pipeline on the GPU, through the C ABI: fpg_resize_bicubic_aa / fpg_tile_gather
and the device-resident FloodDataset / DeviceLoader against the numpy oracle (oracle/data_oracle.py) and the golden
vectors of the reference's apply_transformations. Tolerance 2e-6 absolute on values in [-1, 1]: float32 resampling sums,
the kernels keep the reference's summation order (tap order, horizontal pass first) without fused multiply-adds."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_data as G  # noqa: E402  (seeded synthetic "decoded TIFFs"; the reference is not imported)

pytestmark = pytest.mark.gpu
VECTORS = json.load(open(os.path.join(HERE, "golden", "data_vectors.json")))
FIXTURE = os.path.join(HERE, "golden", "dataset_split_fixture.csv")
TOL = 2e-6


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def device_transform(x_hwc, y_hwc, topography, resize, crop, crop_index, flipped):
    """the product path for one sample: resize kernels + gather kernel"""
    from fpgan import ops
    from models import data
    outs = []
    for host, cmap in ((x_hwc, data.TOPOGRAPHY_CHANNEL_MAP[topography]), (y_hwc, (0, 1, 2))):
        dev = torch.from_numpy(host).cuda()
        oh, ow = data.resize_output_size(host.shape[0], host.shape[1], resize)
        img = ops.resize_bicubic_aa(dev, cmap, oh, ow, flip_w=flipped)
        div = int(np.sqrt(crop)) if crop else 1
        ptrs = torch.tensor([img.data_ptr()], dtype=torch.int64, device="cuda")
        crops = torch.tensor([crop_index if crop else 0], dtype=torch.int32, device="cuda")
        out = torch.empty(1, len(cmap), oh // div, ow // div, device="cuda")
        ops.tile_gather(ptrs, crops, len(cmap), oh, ow, div, out)
        outs.append(out[0].cpu().numpy())
    return outs


@pytest.mark.parametrize("case", VECTORS["transforms"], ids=lambda c: f"seed{c['seed']}")
def test_device_transform_matches_oracle_and_reference(case):
    from oracle import data_oracle as DO
    x, y = G.decoded_pair(case["seed"], case["h"], case["w"])
    args = (case["topography"], case["resize"], case["crop"], case["crop_index"], case["flipped"])
    got = device_transform(x, y, *args)
    want = DO.apply_transformations(x, y, *args)
    for g, w, ref in zip(got, want, (case["input"], case["output"])):
        assert g.shape == w.shape
        assert np.abs(g - w).max() <= TOL
        f = torch.from_numpy(g).double().reshape(-1)  # and directly against the reference's own output
        idx = torch.linspace(0, f.numel() - 1, len(ref["samples"])).long()
        assert (f[idx] - torch.tensor(ref["samples"], dtype=torch.float64)).abs().max().item() <= TOL


def test_resize_properties_full_size():
    """size-independent properties at the training size (1024 -> 512, 9 channels): constants are preserved, the
    operator is linear, and a flipped source gives the mirrored result"""
    from fpgan import ops
    g = torch.Generator().manual_seed(3)
    a = torch.rand(1024, 1024, 9, generator=g).cuda()
    b = torch.rand(1024, 1024, 9, generator=g).cuda()
    cmap = tuple(range(9))
    ra, rb = ops.resize_bicubic_aa(a, cmap, 512, 512), ops.resize_bicubic_aa(b, cmap, 512, 512)
    rab = ops.resize_bicubic_aa(a + 2 * b, cmap, 512, 512)
    assert (rab - (ra + 2 * rb)).abs().max().item() < 5e-6
    const = ops.resize_bicubic_aa(torch.full((1024, 1024, 9), 0.375, device="cuda"), cmap, 512, 512)
    assert (const - 0.375).abs().max().item() < 1e-6
    flipped = ops.resize_bicubic_aa(a, cmap, 512, 512, flip_w=True)
    assert (flipped - ra.flip(-1)).abs().max().item() < 2e-6
    sel = ops.resize_bicubic_aa(a, (0, 1, 2, 5), 512, 512)
    assert torch.equal(sel, ra[[0, 1, 2, 5]])


def test_resize_rejects_bad_arguments():
    from fpgan import ops
    a = torch.rand(8, 8, 3, device="cuda")
    with pytest.raises(RuntimeError):
        ops.resize_bicubic_aa(a, (0, 1, 7), 4, 4)  # channel outside the stack


class _Decoder:
    """stands in for tifffile.imread: a seeded HWC float32 stack per path, counting how often a file is decoded"""

    def __init__(self, size):
        self.size, self.calls = size, {}

    def __call__(self, path):
        self.calls[path] = self.calls.get(path, 0) + 1
        seed = int.from_bytes(path.encode()[-6:], "little") % (2 ** 31)
        c = 9 if "dataset_input" in path else 3
        return np.random.RandomState(seed).rand(self.size, self.size, c).astype(np.float32)


def test_loader_matches_oracle_and_decodes_once():
    from models import data
    from oracle import data_oracle as DO
    dec = _Decoder(64)
    train, val, test = data.create_flood_dataset("harveyonflorence", "best", "/data", "map", 32, 4, batch_size=5,
                                                 decoder=dec, metadata_csv=FIXTURE)
    files = data.determine_flood_dataset("harveyonflorence", "best", 4, FIXTURE)
    assert len(train.dataset) == len(files["train"]) and len(val.dataset) == len(files["validation"])
    assert any(v == "flipped" for _, v, _ in files["train"])
    torch.manual_seed(11)
    order = data.DeviceLoader(train.dataset, 5).order()
    torch.manual_seed(11)
    seen = 0
    for bi, (x, y, names) in enumerate(train):
        idx = order[bi * 5:(bi + 1) * 5]
        assert x.shape == (len(idx), 6, 16, 16) and y.shape == (len(idx), 3, 16, 16) and x.is_cuda
        for k, i in enumerate(idx):
            fname, version, crop = files["train"][i]
            xin = dec(f"/data/dataset_input/{fname}")
            yin = dec(f"/data/dataset_output/{fname[:-8]}.tif")
            wx, wy = DO.apply_transformations(xin, yin, "map", 32, 4, int(crop), version == "flipped")
            assert np.abs(x[k].cpu().numpy() - wx).max() <= TOL and np.abs(y[k].cpu().numpy() - wy).max() <= TOL
            assert names[k] == f"{fname[:-8]}_{crop}"
        seen += len(idx)
        if bi == 3:
            break
    assert seen == 20
    # every file was decoded once by the store (+ the oracle's own calls above), however many crops were drawn
    store = train.dataset.store
    assert store.bytes() == len(store.images) * (6 + 3) * 32 * 32 * 4
    x0, y0, name0 = train.dataset[order[0]]
    assert x0.shape == (6, 16, 16) and name0.endswith(f"_{files['train'][order[0]][2]}")


def test_loader_without_crop_or_resize():
    from models import data
    from oracle import data_oracle as DO
    dec = _Decoder(24)
    _, _, test = data.create_flood_dataset("usa", "same", "/d", None, None, None, batch_size=2, decoder=dec,
                                           metadata_csv=FIXTURE)
    files = data.determine_flood_dataset("usa", "same", None, FIXTURE)["test"]
    torch.manual_seed(2)
    order = data.DeviceLoader(test.dataset, 2).order()
    torch.manual_seed(2)
    x, y, names = next(iter(test))
    assert x.shape == (2, 3, 24, 24)
    for k, i in enumerate(order[:2]):
        fname, version = files[i]
        wx, wy = DO.apply_transformations(dec(f"/d/dataset_input/{fname}"), dec(f"/d/dataset_output/{fname[:-8]}.tif"),
                                          None, None, None, 0, version == "flipped")
        assert np.abs(x[k].cpu().numpy() - wx).max() <= TOL and np.abs(y[k].cpu().numpy() - wy).max() <= TOL
        assert names[k] == fname[:-8]


def test_model_trains_from_the_device_loader():
    """the loaders plug into Model.train_paired (reference model.py:598-658) as the reference's DataLoaders do: one epoch
    over a small split, one loss entry per epoch, every batch consumed in DataLoader order"""
    from models import data
    from models import model as M
    dec = _Decoder(128)
    train, _, _ = data.create_flood_dataset("midwest-flooding", "best", "/data", "all", 128, 4, batch_size=4, decoder=dec,
                                            metadata_csv=FIXTURE)
    assert len(train) >= 3
    m = M.Model(model="PairedAttention", topography="all", num_epochs=1, seed=47, resize=128, crop=4)
    m.train_loader = train
    m.train_paired()
    for key, values in m.all_losses.items():
        assert len(values) == 1 and np.isfinite(values[0]), (key, values)
    # every (file, version) of the split was decoded exactly once per tensor, whatever the number of crops drawn
    # (training images exist as an original and a flipped version: at most two decodes per file)
    keys = {(f, v) for f, v, _ in train.dataset.data_files}
    assert set(train.dataset.store.images) == keys
    assert sum(dec.calls.values()) == 2 * len(keys) and max(dec.calls.values()) <= 2
