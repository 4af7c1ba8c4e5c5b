"""Generates tests/golden/data_vectors.json (+ the small metadata fixture dataset_split_fixture.csv) by running the
UNMODIFIED reference input pipeline -- models/utils.py:apply_transformations and models/data.py:determine_flood_dataset,
imported from /root/reference, which exists only in the authoring container -- on seeded synthetic "decoded TIFFs".

    python tests/golden/make_golden_data.py

tifffile is not installed: an empty stub module is injected for the import (no image is read from disk here; the arrays
that tifffile.imread would return are synthesised). The fixture CSV has the reference metadata file's columns and value
sets but made-up image names; the reference function is run with the fixture as its metadata/dataset_split.csv.
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def write_fixture_csv(path):
    rng = np.random.RandomState(7)
    rows = []
    plan = [("hurricane-harvey", "usa", 70, ["01m", "10m"], "10m"), ("hurricane-florence", "usa", 30, ["10m", "01m"], "10m"),
            ("midwest-flooding", "usa", 24, ["10m"], "10m"), ("nepal-flooding", "india", 40, ["30m"], "30m")]
    for disaster, country, n, best, same in plan:
        for i in range(n):
            name = f"{disaster}_{int(rng.randint(0, 600)):08d}x{i}"
            split = rng.choice(["train", "train", "train", "train", "train", "train", "train", "validation", "test"])
            b = best[int(rng.randint(0, len(best)))]
            rows.append((name, b, same, "original", split, disaster, country))
    for r in list(rows):  # training images also exist as a flipped version, as in the reference metadata
        if r[4] == "train":
            rows.append(r[:3] + ("flipped",) + r[4:])
    with open(path, "w") as f:
        f.write("image,best_DEM,same_DEM,version,split,disaster,country\n")
        for r in rows:
            f.write(",".join(r) + "\n")


def import_reference():
    sys.modules["tifffile"] = types.ModuleType("tifffile")
    sys.path.insert(0, REF)
    from models import data as ref_data
    from models import utils as ref_utils
    return ref_data, ref_utils


def decoded_pair(seed, h, w):
    """what tifffile.imread returns for a (dataset_input, dataset_output) pair: HWC float32 in [0, 1]"""
    g = np.random.RandomState(seed)
    return g.rand(h, w, 9).astype(np.float32), g.rand(h, w, 3).astype(np.float32)


def digest(t, n=48):
    f = t.double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return {"shape": list(t.shape), "samples": f[idx].tolist(), "sum": f.sum().item(), "abs_sum": f.abs().sum().item()}


def transform_cases(ref_utils):
    cases = []
    grid = [(11, 64, 64, "all", 32, 4, 3, False), (12, 64, 64, "dem", 32, 4, 0, False),
            (13, 64, 64, "flow", 32, 4, 1, True), (14, 64, 64, "river", 32, None, 0, False),
            (15, 64, 64, "map", None, 4, 2, False), (16, 64, 64, None, 32, 4, 2, True),
            (17, 96, 96, "all", 40, 4, 1, False),      # scale 2.4: fractional tap positions
            (18, 48, 80, "all", 32, None, 0, False),   # non-square: smaller edge -> 32, the other keeps the ratio
            (19, 40, 40, "all", 64, 16, 5, False),     # upscale (support 2, no anti-aliasing) and a 4 x 4 crop grid
            (20, 1024, 1024, "all", 512, 4, 2, False)]  # the training configuration (resize=512, crop=4)
    for seed, h, w, topo, resize, crop, crop_index, flipped in grid:
        x, y = decoded_pair(seed, h, w)
        if flipped:  # models/data.py:63-65
            xi = torch.from_numpy(np.fliplr(x).transpose(2, 0, 1).copy())
            yi = torch.from_numpy(np.fliplr(y).transpose(2, 0, 1).copy())
        else:
            xi = torch.from_numpy(x.transpose(2, 0, 1))
            yi = torch.from_numpy(y.transpose(2, 0, 1))
        a, b, name = ref_utils.apply_transformations(image_name="img", input_image=xi, output_image=yi, topography=topo,
                                                     resize=resize, crop=crop, to_loader=True, crop_index=crop_index)
        cases.append({"seed": seed, "h": h, "w": w, "topography": topo, "resize": resize, "crop": crop,
                      "crop_index": crop_index, "flipped": flipped, "name": name, "input": digest(a),
                      "output": digest(b)})
    return cases


def split_cases(ref_data, fixture):
    out = []
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "metadata"))
    shutil.copy(fixture, os.path.join(tmp, "metadata", "dataset_split.csv"))
    cwd = os.getcwd()
    os.chdir(tmp)  # the reference reads metadata/dataset_split.csv relative to the cwd (models/data.py:90)
    try:
        for subset in ["all", "usa", "India", "hurricane-harvey", "nepal-flooding", "harveyflorence", "harveyonflorence",
                       "testing"]:
            for dem in ["best", "same"]:
                for crop in [None, 4]:
                    splits = ref_data.determine_flood_dataset(subset, dem, crop)
                    rec = {"subset": subset, "dem": dem, "crop": crop}
                    for k, items in splits.items():
                        items = [tuple(int(v) if isinstance(v, (int, np.integer)) else v for v in it) for it in items]
                        rec[k] = {"n": len(items), "head": [list(it) for it in items[:3]],
                                  "sha1": hashlib.sha1(repr(items).encode()).hexdigest()}
                    out.append(rec)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp)
    return out


def history_buffer_case():
    """get_buffer_image (models/model.py:275-294) over 160 generated images, Python RNG seeded with 5: which image each
    call returns and what the 50-entry buffer holds at the end (images are tagged with their index)"""
    import random
    import make_golden
    ref_model = make_golden.import_reference()
    buf, returned = [], []
    random.seed(5)
    for i in range(160):
        out = ref_model.Model.get_buffer_image(None, torch.full((1, 1, 2, 2), float(i)), buf)
        returned.append(int(out.flatten()[0].item()))
    return {"py_seed": 5, "n": 160, "returned": returned, "final_buffer": [int(b.flatten()[0].item()) for b in buf]}


def bare_model(model_cls, **attrs):
    """an object with the given attributes on which the (unbound) helper methods of a Model class can be called without
    constructing networks"""
    class Bare:
        pass
    obj = Bare()
    for k, v in attrs.items():
        setattr(obj, k, v)
    obj.prettify_model_name = types.MethodType(model_cls.prettify_model_name, obj)
    return obj


PATH_CASES = [dict(model="pairedattention", model_is_cycle=False, add_identity_loss=False, data_path="/d", current_epoch=7,
                   training_model=True, topography="all", dataset_subset="usa", dataset_dem="best", resize=512, crop=4,
                   save_type="model", info=""),
              dict(model="cyclegan", model_is_cycle=True, add_identity_loss=True, data_path="C:/data", current_epoch=3,
                   training_model=False, topography=None, dataset_subset="all", dataset_dem="same", resize=None, crop=None,
                   save_type="metric", info="val"),
              dict(model="pix2pix", model_is_cycle=False, add_identity_loss=False, data_path="rel", current_epoch=1,
                   training_model=True, topography="dem", dataset_subset="india", dataset_dem="best", resize=256, crop=None,
                   save_type="image", info="sample_3")]


def mask_date(path):
    import re
    return re.sub(r"date\d{4}-\d{2}-\d{2}-\d{2}-\d{2}-\d{2}", "date<T>", path)


def model_helper_cases():
    """lambda_rule, initialise_loss_storage, create_path (date masked) of the reference's Model (model.py:175-260)"""
    import make_golden
    cls = make_golden.import_reference().Model
    out = {"lambda_rule": {}, "loss_keys": [], "paths": []}
    for n in (1, 2, 7, 200):
        out["lambda_rule"][str(n)] = [cls.lambda_rule(bare_model(cls, num_epochs=n), e) for e in range(n + 2)]
    for cycle in (False, True):
        for identity in (False, True):
            for overall in (False, True):
                keys = list(cls.initialise_loss_storage(bare_model(cls, model_is_cycle=cycle, add_identity_loss=identity),
                                                        overall).keys())
                out["loss_keys"].append({"cycle": cycle, "identity": identity, "overall": overall, "keys": keys})
    for case in PATH_CASES:
        attrs = {k: v for k, v in case.items() if k not in ("save_type", "info")}
        out["paths"].append(mask_date(cls.create_path(bare_model(cls, **attrs), case["save_type"], case["info"])))
    return out


def main():
    fixture = os.path.join(HERE, "dataset_split_fixture.csv")
    write_fixture_csv(fixture)
    ref_data, ref_utils = import_reference()
    import torchvision
    vectors = {"torch": torch.__version__, "torchvision": torchvision.__version__,
               "transforms": transform_cases(ref_utils), "splits": split_cases(ref_data, fixture)}
    vectors["history_buffer"] = history_buffer_case()  # last: importing the reference's model module changes the cwd
    vectors["model_helpers"] = model_helper_cases()
    with open(os.path.join(HERE, "data_vectors.json"), "w") as f:
        json.dump(vectors, f)
    print("wrote", len(vectors["transforms"]), "transform cases,", len(vectors["splits"]), "split cases")


if __name__ == "__main__":
    main()
