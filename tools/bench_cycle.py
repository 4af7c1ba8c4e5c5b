"""Throughput of the cycle-consistent training step (BASELINE configs[3]: CycleGAN and AttentionGAN, reference
Model.train_cycle, models/model.py:660-758) and of the Pix2Pix paired step through the drop-in module path: every network
call is a native executor, the losses / optimisers / history buffer are the reference's torch host code.
Usage: python tools/bench_cycle.py [batch] [steps]   -> one JSON line per model (tiles/s, wall clock with a final sync)."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
import torch  # noqa: E402

from models import model as M  # noqa: E402
from models.data import SyntheticLoader  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for name, identity in (("CycleGAN", False), ("AttentionGAN", False), ("AttentionGAN", True), ("Pix2Pix", False)):
    m = M.Model(model=name, topography="all", num_epochs=200, seed=47, add_identity_loss=identity)
    run = m.train_cycle if m.model_is_cycle else m.train_paired
    warm = list(SyntheticLoader(steps=3, batch=batch))
    data = list(SyntheticLoader(steps=steps, batch=batch))
    m.train_loader = warm
    m.starting_epoch = m.num_epochs
    m.all_losses = m.initialise_loss_storage(overall=True)
    run()
    torch.cuda.synchronize()
    m.train_loader = data
    m.starting_epoch = m.num_epochs
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"model": name, "identity_loss": identity, "batch": batch, "steps": steps,
                      "ms_per_step": 1e3 * dt / steps, "tiles_per_s": batch * steps / dt,
                      "path": "drop-in modules (native executors) + reference host loop, pinned host batches"}),
          flush=True)
