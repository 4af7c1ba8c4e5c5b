"""Same-box GPU comparator: the oracle restatement of Model.train_paired, i.e. the reference's own
sequence of PyTorch operators, run by eager PyTorch + cuDNN on the B200 -- what a user gets by moving the unmodified
reference to this GPU. Not a test and not part of the product: it lives under tests/ because only tests may execute
oracle/. Usage (on a GPU box): python tests/gpu_comparator.py [batch] [steps]  -> one JSON line per precision mode."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gan_oracle as O  # noqa: E402


def run(mode, batch, steps, warmup=3):
    torch.backends.cudnn.benchmark = True
    tf32 = mode != "fp32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    nets = O.init_model("pairedattention", "all", seed=47)
    for net in nets.values():
        for k in list(net):
            net[k] = net[k].cuda()
    tr = O.PairedTrainer(nets)
    data = [tuple(t.cuda() for t in O.synthetic_batch(s, batch)) for s in range(4)]
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else torch.autocast("cuda", enabled=False)
    torch.set_default_device("cuda")  # the restatement builds its LSGAN targets with torch.full(shape, value)
    with ctx:
        for s in range(warmup):
            tr.step(*data[s % 4])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(steps):
            out = tr.step(*data[s % 4])
        b.record()
        torch.cuda.synchronize()
    torch.set_default_device("cpu")
    ms = a.elapsed_time(b) / steps
    return {"comparator": "eager PyTorch + cuDNN, oracle restatement of train_paired", "mode": mode, "batch": batch,
            "steps": steps, "ms_per_step": ms, "tiles_per_s": batch * 1000.0 / ms,
            "loss_l1_last": out["l1_losses_generator_synthetic"], "torch": torch.__version__,
            "gpu": torch.cuda.get_device_name(0)}


if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    for mode in ("fp32", "tf32", "bf16_autocast"):
        print(json.dumps(run(mode, batch, steps)), flush=True)
