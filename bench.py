#!/usr/bin/env python
"""Benchmark of the hot path: the PairedAttention `train_paired` step (reference models/model.py:611-651) on
synthetic 256x256 tiles (resize=512, crop=4, topography=all -> 9 input channels), batch 16 per GPU, bf16 tensor-core
compute with fp32 accumulation.

    python bench.py --gpus N --steps K --warmup W            # this implementation (N>1: launched by torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

Prints ONE JSON line (rank 0). See DESIGN.md section "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "flood-prediction-gan_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_RESULT_OUT = sys.stdout
TILE = 256
CHANNELS = 9
BATCH_PER_GPU = 16
GFLOP_PER_TILE = 397.31           # algorithmic FLOPs of one paired step per tile
RES_CONV_GFLOP_PER_TILE = 4.8318  # one residual 3x3 conv, 256->256 @ 64x64: 2 * 4096 * 256 * 2304


def ncu_traffic_bytes(report="r01_res_fprop.ncu-rep"):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed ncu --set full summary
    (profiles/r01_ncu_kernels_v9.csv, produced by tools/profile_kernels.sh + tools/summarize_ncu.py); None if absent."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_kernels_v9.csv")
    if not os.path.exists(path):
        return None
    import csv
    for row in csv.DictReader(open(path)):
        if row["report"] == report:
            return int((float(row["dram_read_MB"]) + float(row["dram_write_MB"])) * 1e6)
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1413.9), p.get("hbm_gbs", 6464.9), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons (NVML, every 20 ms) while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.max_mhz = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self.stop.is_set():
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((mhz, [n for n, b in bits.items() if mask & b]))
                self.stop.wait(0.02)
        except Exception:
            self._run_smi()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6 and f[0].isdigit():
                    self.max_mhz = int(f[1])
                    self.samples.append((int(f[0]), [n for i, n in enumerate(names)
                                                     if f[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[1]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_reference_tiles_per_s(steps, warmup, batch=1):
    """The reference algorithm on the host CPU: oracle port of Model.train_paired, fp32, all host threads."""
    from oracle import gan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets = O.init_model("pairedattention", "all", seed=47)
    tr = O.PairedTrainer(nets)
    for s in range(warmup):
        tr.step(*O.synthetic_batch(s, batch, CHANNELS, TILE))
    data = [O.synthetic_batch(warmup + s, batch, CHANNELS, TILE) for s in range(steps)]
    t0 = time.perf_counter()
    for x, y in data:
        tr.step(x, y)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    tps, dt, cores = cpu_reference_tiles_per_s(steps, warmup, batch=1)
    sample = (f"oracle port of train_paired (oracle/gan_oracle.py), fp32, batch 1 of the batch-16 workload per step, "
              f"{steps} timed steps after {warmup} warm-up")
    line = {"impl": "reference", "metric": "PairedAttention train tiles/s", "value": tps, "unit": "tiles/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1000 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "PairedAttention train_paired step, 256x256 tiles (resize=512 crop=4), 9 input "
                                   "channels (topography=all), batch 16 per GPU"},
            "cpu_baseline": {"value": tps, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def run_native(args):
    import torch.distributed as dist
    from fpgan import ops
    from fpgan.trainer import PairedTrainer
    from models import model as M
    from models.data import SyntheticLoader

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this implementation has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("FPG_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, steps, warmup = args.batch, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    model = M.Model(model="PairedAttention", topography="all", num_epochs=1, resize=512, crop=4, seed=47,
                    log_interval=1)
    tr = model._ensure_native_paired()

    # ---- (1) device-resident inputs: `value`
    loader = SyntheticLoader(steps=4, batch=B, channels=CHANNELS, size=TILE, rank=rank, world_size=world, pin=False)
    resident = [(x.to(dev), y.to(dev)) for x, y, _ in loader]
    for s in range(warmup):
        tr.step(*resident[s % len(resident)])
    barrier()
    ops.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for s in range(steps):
            tr.step(*resident[s % len(resident)])
        e1.record()
        barrier()
    launches = ops.LAUNCHES
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = B * world * steps / (ms / 1000.0)
    last_losses = tr.losses()

    # ---- (2) end to end through the public API: Model.train_paired with pinned HOST batches; every step copies its
    # inputs host->device and reads the four losses back (log_interval=1)
    host = list(SyntheticLoader(steps=steps, batch=B, channels=CHANNELS, size=TILE, rank=rank, world_size=world))
    model.train_loader = host[:min(3, steps)]
    model.num_epochs = model.starting_epoch = 1
    model.train_paired()  # warm-up epoch of the API path
    model.train_loader = host
    model.starting_epoch = model.num_epochs = 2
    barrier()
    e0.record()
    t0 = time.perf_counter()
    model.train_paired()
    e1.record()
    barrier()
    wall_ms = 1000 * (time.perf_counter() - t0)
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    e2e = B * world * steps / (e2e_ms / 1000.0)

    # ---- (3) roofline of the dominant kernel: igemm_fprop_kernel<64> on the residual 3x3 conv (18 launches per
    # generator forward, 75% of the generator FLOPs), timed with CUDA events around each launch of extra steps
    ops.PROFILE = {}
    for s in range(2):
        tr.step(*resident[s % len(resident)])
    torch.cuda.synchronize()
    prof = {}
    for key, evs in ops.PROFILE.items():
        prof[key] = [a.elapsed_time(b) for a, b in evs]
    ops.PROFILE = None
    peak_tf, peak_hbm, peak_src = measured_peaks()
    res_key = f"fprop n{B} 64x64 c256 k256 r3 s1"
    res_ms = prof.get(res_key, [])
    roofline = None
    if res_ms:
        avg = sum(res_ms) / len(res_ms)
        achieved = RES_CONV_GFLOP_PER_TILE * B / avg  # GFLOP / ms == TFLOP/s
        roofline = {"bound": "tensor", "kernel": "igemm_fprop2_kernel<64> (2-CTA tcgen05, residual 3x3 conv 256->256 @64x64)",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": ncu_traffic_bytes(), "peak_source": peak_src + ", sustained bf16", "launches_timed": len(res_ms),
                    "avg_ms": avg}
    conv_ms = sum(sum(v) for k, v in prof.items() if k.split()[0] in ("fprop", "dgrad", "wgrad")) / 2
    all_ms = sum(sum(v) for v in prof.values()) / 2
    step_tflops = GFLOP_PER_TILE * B * world / (ms / steps)

    if rank == 0:
        cpu_tps, cpu_dt, cores = cpu_reference_tiles_per_s(steps=8, warmup=1, batch=1) if world == 1 else (None, 0, 0)
        line = {"metric": "PairedAttention train tiles/s", "value": value, "unit": "tiles/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "PairedAttention train_paired step, 256x256 tiles (resize=512 crop=4), 9 input "
                                       "channels (topography=all), batch 16 per GPU", "batch_per_gpu": B,
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "per-step working set (>5 GB of activations) far exceeds the 126 MB L2"},
                "clocks": clocks.summary(),
                "e2e": {"value": e2e, "unit": "tiles/s", "h2d_bytes_per_step": B * (CHANNELS + 3) * TILE * TILE * 4,
                        "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / steps,
                        "api": "models.model.Model.train_paired() with pinned host batches, log_interval=1"},
                "gpu_launches": launches,
                "roofline": roofline,
                "step_tflops": step_tflops, "step_frac_of_peak": step_tflops / (peak_tf * world),
                "conv_share_of_step": conv_ms / all_ms if all_ms else None,
                "kernel_ms_per_step": {k: sum(v) / 2 for k, v in sorted(prof.items(), key=lambda kv: -sum(kv[1]))[:12]},
                "losses_last_step": last_losses}
        if cpu_tps is not None:
            line["cpu_baseline"] = {"value": cpu_tps, "unit": "tiles/s", "cores": cores, "kind": "port",
                                    "sample": "oracle port of train_paired, fp32, 8 timed batch-1 steps (256x256) "
                                              "after 1 warm-up"}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        tr.release_graphs()  # captured NCCL work must be gone before the communicator is torn down
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: libraries that print to file descriptor 1 (NCCL prints its version there)
    # are diverted to stderr for the whole run and the result is written to the saved descriptor
    global _RESULT_OUT
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
