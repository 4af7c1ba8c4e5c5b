"""Data-parallel parity ON HARDWARE (SURVEY.md section 4 "Distributed tests"): N ranks x B/N tiles must reproduce one
process x B tiles. Launches `bench.py --check` under torchrun on 2 GPUs; skipped on a single-GPU box (the result of the
2- and 8-GPU runs made during development is committed under profiles/)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["pairedattention", "cyclegan"])
def test_two_ranks_match_one_process(model):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--check", "--model",
           model, "--batch", "4", "--check_size", "128"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and lines, out.stderr[-3000:]
    rec = json.loads(lines[-1])
    print(rec)
    assert rec["ok"], rec
