"""Input pipeline, CPU side: the numpy oracle of apply_transformations against the golden
vectors produced by the reference's own code (tests/golden/make_golden_data.py), the dataset-split logic against the
reference's lists on the fixture metadata, and the loader order against torch's DataLoader(shuffle=True)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle import data_oracle as DO  # noqa: E402
from models import data  # noqa: E402
import make_golden_data as G  # noqa: E402  (only its seeded synthetic-image helper is used; the reference is not imported)

VECTORS = json.load(open(os.path.join(HERE, "golden", "data_vectors.json")))
FIXTURE = os.path.join(HERE, "golden", "dataset_split_fixture.csv")
TOL = 2e-6  # float32 resampling: the restatement agrees with ATen to 1-4 ulp of values in [-1, 1]


def check_digest(got, ref):
    assert list(got.shape) == ref["shape"]
    f = torch.from_numpy(np.ascontiguousarray(got)).double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, len(ref["samples"])).long()
    assert (f[idx] - torch.tensor(ref["samples"], dtype=torch.float64)).abs().max().item() <= TOL
    assert abs(f.sum().item() - ref["sum"]) <= TOL * f.numel() ** 0.5 * 4
    assert abs(f.abs().sum().item() - ref["abs_sum"]) <= TOL * f.numel() ** 0.5 * 4


@pytest.mark.parametrize("case", VECTORS["transforms"], ids=lambda c: f"seed{c['seed']}")
def test_oracle_matches_reference_transformations(case):
    x, y = G.decoded_pair(case["seed"], case["h"], case["w"])
    a, b = DO.apply_transformations(x, y, case["topography"], case["resize"], case["crop"], case["crop_index"],
                                    case["flipped"])
    check_digest(a, case["input"])
    check_digest(b, case["output"])


def test_resampling_weights_properties():
    for n_in, n_out in ((1024, 512), (96, 40), (40, 64), (7, 3)):
        table = DO.aa_weights(n_in, n_out)
        assert len(table) == n_out
        for lo, n, w in table:
            assert 0 <= lo and lo + n <= n_in and n >= 1
            assert abs(float(w.sum()) - 1.0) < 1e-6  # a constant image stays constant
    img = np.full((2, 12, 20), 0.25, dtype=np.float32)
    assert np.allclose(DO.resize_bicubic_aa(img, 6), 0.25, atol=1e-6)
    assert DO.resize_output_size(48, 80, 32) == (32, 53) and DO.resize_output_size(80, 48, 32) == (53, 32)
    assert data.resize_output_size(48, 80, 32) == (32, 53) and data.resize_output_size(64, 64, None) == (64, 64)


@pytest.mark.parametrize("rec", VECTORS["splits"], ids=lambda r: f"{r['subset']}-{r['dem']}-{r['crop']}")
def test_dataset_split_matches_reference(rec):
    got = data.determine_flood_dataset(rec["subset"], rec["dem"], rec["crop"], metadata_csv=FIXTURE)
    for split in ("train", "validation", "test"):
        items = [tuple(int(v) if isinstance(v, (int, np.integer)) else v for v in it) for it in got[split]]
        assert len(items) == rec[split]["n"]
        assert [list(it) for it in items[:3]] == rec[split]["head"]
        assert hashlib.sha1(repr(items).encode()).hexdigest() == rec[split]["sha1"]


def test_dataset_split_rejects_unknown_names():
    with pytest.raises(NotImplementedError):
        data.determine_flood_dataset("atlantis", "best", None, metadata_csv=FIXTURE)
    with pytest.raises(NotImplementedError):
        data.determine_flood_dataset("usa", "worst", None, metadata_csv=FIXTURE)


class _Indices(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


@pytest.mark.parametrize("n,batch", [(10, 4), (37, 1), (64, 16)])
def test_loader_order_is_torch_dataloader_order(n, batch):
    """same global seed -> the same sample order as DataLoader(shuffle=True) (data.py:28-32), epoch after epoch"""
    torch.manual_seed(47)
    ref_loader = torch.utils.data.DataLoader(_Indices(n), batch_size=batch, shuffle=True, num_workers=0)
    ref = [[b.tolist() for b in ref_loader] for _ in range(2)]
    torch.manual_seed(47)
    mine = data.DeviceLoader(_Indices(n), batch_size=batch, shuffle=True)
    for epoch in range(2):
        order = mine.order()
        got = [order[s:s + batch] for s in range(0, n, batch)]
        assert got == ref[epoch]
    assert len(mine) == len(ref_loader)


def test_sharded_loader_partitions_the_global_batch():
    n, batch, world = 23, 3, 2

    class Rec(_Indices):
        def gather(self, idx):
            return list(idx)

    per_rank = []
    for rank in range(world):
        torch.manual_seed(5)
        per_rank.append(list(data.DeviceLoader(Rec(n), batch, rank=rank, world_size=world)))
    torch.manual_seed(5)
    single = list(data.DeviceLoader(Rec(n), batch * world))
    # every rank takes the same number of equally sized batches (the fused step's NCCL all-reduces need every rank):
    # 23 samples over 2 ranks x 3 -> 4 global batches, the ragged last one completed by wrapping around
    assert len(per_rank[0]) == len(per_rank[1]) == len(data.DeviceLoader(Rec(n), batch, world_size=world)) == 4
    assert all(len(b) == batch for r in per_rank for b in r)
    merged = [a + b for a, b in zip(per_rank[0], per_rank[1])]
    assert merged[:-1] == single[:-1]
    torch.manual_seed(5)
    order = data.DeviceLoader(Rec(n), batch * world).order()
    assert merged[-1] == single[-1] + order[:1]


@pytest.mark.parametrize("n,batch,world", [(5, 4, 2), (7, 2, 4), (8, 2, 2), (1, 2, 2)])
def test_sharded_loader_equal_step_counts(n, batch, world):
    """dataset lengths that do not divide world * batch (and one shorter than a single global batch)"""
    class Rec(_Indices):
        def gather(self, idx):
            return list(idx)

    shards = []
    for rank in range(world):
        torch.manual_seed(9)
        shards.append(list(data.DeviceLoader(Rec(n), batch, rank=rank, world_size=world)))
    assert len({len(s) for s in shards}) == 1 and len(shards[0]) == -(-n // (batch * world))
    assert all(len(b) == batch for s in shards for b in s)
    seen = {i for s in shards for b in s for i in b}
    assert seen == set(range(n))  # nothing is dropped


def test_create_flood_dataset_without_data_returns_empty_loaders():
    assert data.create_flood_dataset("all", "best", None, "all", 256, None) == ([], [], [])


def test_history_buffer_matches_reference():
    """get_buffer_image (reference model.py:275-294) beyond the 50-entry fill phase: the product's method (global Python
    RNG, as the reference) and the oracle's restatement return the reference's images and end with its buffer"""
    import random

    from oracle import gan_oracle as GO
    from models import model as M
    gold = VECTORS["history_buffer"]

    def run(fn):
        buf, out = [], []
        for i in range(gold["n"]):
            out.append(int(fn(torch.full((1, 1, 2, 2), float(i)), buf).flatten()[0].item()))
        return out, [int(b.flatten()[0].item()) for b in buf]

    random.seed(gold["py_seed"])
    got, final = run(lambda img, buf: M.Model.get_buffer_image(None, img, buf))
    assert got == gold["returned"] and final == gold["final_buffer"]
    tr = GO.CycleTrainer.__new__(GO.CycleTrainer)
    tr.rng = random.Random(gold["py_seed"])
    got, final = run(tr._buffer)
    assert got == gold["returned"] and final == gold["final_buffer"]
    assert any(r != i for i, r in enumerate(gold["returned"]))  # the replacement branch was exercised
    # the fused cycle step keeps the pool on the device and takes only the DECISIONS on the host
    # (fpgan.trainer._History -> {use_slot, store_slot} for fpg_history_exchange): replay them on a host pool
    from fpgan.trainer import _History
    random.seed(gold["py_seed"])
    hist, pool, got = _History(), [None] * _History.SIZE, []
    for i in range(gold["n"]):
        use, store = hist.decide()
        got.append(pool[use] if use >= 0 else i)
        if store >= 0:
            pool[store] = i
    assert got == gold["returned"] and pool == gold["final_buffer"]


def test_model_helpers_match_reference():
    """lambda_rule (learning-rate schedule), initialise_loss_storage (loss-dictionary keys and order) and create_path
    (artefact naming, date masked) of Model against the reference's own methods (model.py:175-260); the oracle's
    lambda_rule too"""
    from oracle import gan_oracle as GO
    from models import model as M
    gold = VECTORS["model_helpers"]
    for n, want in gold["lambda_rule"].items():
        got = [M.Model.lambda_rule(G.bare_model(M.Model, num_epochs=int(n)), e) for e in range(int(n) + 2)]
        assert got == want
        assert [GO.lambda_rule(e, int(n)) for e in range(int(n) + 2)] == want
    for rec in gold["loss_keys"]:
        bare = G.bare_model(M.Model, model_is_cycle=rec["cycle"], add_identity_loss=rec["identity"])
        assert list(M.Model.initialise_loss_storage(bare, rec["overall"]).keys()) == rec["keys"]
    for case, want in zip(G.PATH_CASES, gold["paths"]):
        attrs = {k: v for k, v in case.items() if k not in ("save_type", "info")}
        got = M.Model.create_path(G.bare_model(M.Model, **attrs), case["save_type"], case["info"])
        assert G.mask_date(got) == want


def test_metrics_csv_is_the_reference_dataframe_csv():
    """calculate_metrics writes what the reference's pandas code writes (model.py:419-421)"""
    import io

    import pandas as pd
    from models import model as M
    results = {"PSNR": 20.476618475778388, "SSIM": 0.9455338178579847, "MS-SSIM": float("nan"), "LPIPS": float("nan"),
               "MSE": 0.25, "Accuracy": 0.75, "F1_Flood": 1 / 3, "Inference": 0.012345678901234}
    frame = pd.DataFrame([(k, np.mean([v])) for k, v in results.items()]).set_index(0).transpose()
    buf = io.StringIO()
    frame.to_csv(buf)
    assert M.Model.metrics_csv_text(results) == buf.getvalue()
