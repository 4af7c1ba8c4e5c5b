"""Drives the staged, unmodified reference (oracle/_ref, see build_ref.py) on the host CPU.

TEST / MEASUREMENT INFRASTRUCTURE, not product code (see build_ref.py). Recipe of SURVEY.md appendix D: the packages
the reference imports for plotting / file decoding / image metrics (tifffile, matplotlib, torchmetrics) are absent, so
empty stub modules stand in for them; `Model.train_loader` is replaced by a list of synthetic batches; nothing in the
reference's maths is touched. `Model.__init__` reads metadata/dataset_split.csv relative to the cwd (data.py:87), hence
the chdir while a model is constructed.
"""
import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_STUBS = {
    "tifffile": (), "matplotlib": (), "matplotlib.pyplot": (), "torchmetrics": (),
    "torchmetrics.regression": ("MeanSquaredError",),
    "torchmetrics.image": ("PeakSignalNoiseRatio", "MultiScaleStructuralSimilarityIndexMeasure",
                           "StructuralSimilarityIndexMeasure"),
    "torchmetrics.image.lpip": ("LearnedPerceptualImagePatchSimilarity",),
    "torchmetrics.classification": ("BinaryAccuracy", "BinaryF1Score", "BinaryPrecision", "BinaryRecall"),
}


def available():
    return os.path.isfile(os.path.join(REF_DIR, "models", "model.py"))


@contextlib.contextmanager
def _reference_imports():
    """sys.path / sys.modules arranged so that `models` is the REFERENCE's package, restored afterwards (this
    repository's own drop-in package is also called `models`)."""
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    added = []
    for name, attrs in _STUBS.items():
        if name not in sys.modules:
            try:
                importlib.import_module(name)
                continue
            except Exception:
                pass
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
            added.append(name)
            if "." in name:
                setattr(sys.modules[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], m)
    # the reference's models/ has no __init__.py (namespace package): a regular package of the same name anywhere on
    # sys.path (this repository's drop-in package) would win, so the package object is created explicitly
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF_DIR, "models")]
    sys.modules["models"] = pkg
    cwd = os.getcwd()
    os.chdir(REF_DIR)
    try:
        yield
    finally:
        os.chdir(cwd)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)


def make_model(model="pairedattention", add_identity_loss=False, num_epochs=200, seed=47):
    """the reference's Model (CPU, fp32) with an empty synthetic loader; assign `.train_loader` a list of
    (input [B,9,H,W], target [B,3,H,W], names) tuples and call `.train_paired()` / `.train_cycle()`"""
    if not available():
        raise FileNotFoundError(f"{REF_DIR} is not staged: run `python oracle/build_ref.py` where /root/reference exists")
    with _reference_imports():
        from models import model as ref_model
        # the reference picks cuda when present (model.py:24, utils.py:6, ...); this arm is its CPU path
        for name, mod in list(sys.modules.items()):
            if (name == "models" or name.startswith("models.")) and hasattr(mod, "device"):
                mod.device = "cpu"
        m = ref_model.Model(model=model, dataset_subset="usa", dataset_dem="same", data_path="/tmp/none",
                            num_epochs=num_epochs, topography="all", resize=512, crop=4, training_model=True, seed=seed,
                            add_identity_loss=add_identity_loss)
    m.train_loader = []
    return m


def run_epoch(m, batches, cycle=False):
    """one pass of the reference's own training loop over `batches`; returns the epoch-mean losses it recorded"""
    m.train_loader = batches
    m.starting_epoch = m.num_epochs  # exactly one epoch; lambda_rule keeps lr = 2e-4 while num_epochs is large
    m.all_losses = m.initialise_loss_storage(overall=True)
    (m.train_cycle if cycle else m.train_paired)()
    return {k: float(v[-1]) for k, v in m.all_losses.items()}
