#!/usr/bin/env python
"""Benchmark of the hot path: the PairedAttention `train_paired` step (reference models/model.py:611-651) on
synthetic 256x256 tiles (resize=512, crop=4, topography=all -> 9 input channels), batch 16 per GPU, bf16 tensor-core
compute with fp32 accumulation.

    python bench.py --gpus N --steps K --warmup W            # this implementation (N>1: launched by torchrun)
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference on the host CPU (oracle/_ref)
    python bench.py --model cyclegan|attentiongan [--identity]   # BASELINE configs[3]: the fused train_cycle step
    torchrun ... bench.py --gpus N --check [--model ...]     # data-parallel parity: N ranks x B/N vs one rank x B

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
# algorithmic GFLOP of one training step per tile, by (model, identity loss)
GFLOP_PER_TILE = {("pairedattention", False): 397.31, ("cyclegan", False): 1314.1, ("cyclegan", True): 1916.2,
                  ("attentiongan", False): 1491.5, ("attentiongan", True): 2182.2, ("pix2pix", False): 88.55}
CYCLE_MODELS = ("cyclegan", "attentiongan")
RES_CONV_GFLOP_PER_TILE = 4.8318  # one residual 3x3 conv, 256->256 @ 64x64: 2 * 4096 * 256 * 2304
PRETTY = {"pairedattention": "PairedAttention", "cyclegan": "CycleGAN", "attentiongan": "AttentionGAN",
          "pix2pix": "Pix2Pix"}


def workload_config(model, identity, batch, world):
    """`config` of the JSON line: identical keys and values on the native and the reference arm"""
    if model in ("pairedattention", "pix2pix"):
        what = f"{PRETTY[model]} train_paired step"
    else:
        what = f"{PRETTY[model]} train_cycle step" + (" with identity loss" if identity else "")
    return {"workload": f"{what}, 256x256 tiles (resize=512 crop=4), 9 input channels (topography=all), batch "
                        f"{batch} per GPU", "batch_per_gpu": batch, "global_batch": batch * world,
            "parallelism": f"dp{world}",
            "l2": "per-step working set (>5 GB of activations) far exceeds the 126 MB L2"}


def ncu_traffic_bytes(report="r02_res_fprop.ncu-rep"):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed ncu --set full summary
    (profiles/r02/ncu_kernels.csv, produced by tools/profile_kernels.sh + tools/summarize_ncu.py); None if absent."""
    path = os.path.join(ROOT, "profiles", "r02", "ncu_kernels.csv")
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


def _synthetic_host_batches(first, steps, batch):
    out = []
    for step in range(first, first + steps):
        g = torch.Generator().manual_seed(1000 + step)
        x = torch.rand(batch, CHANNELS, TILE, TILE, generator=g) * 2 - 1
        y = torch.rand(batch, 3, TILE, TILE, generator=g) * 2 - 1
        out.append((x, y, ("synthetic",) * batch))
    return out


def cpu_reference_tiles_per_s(steps, warmup, batch, model="pairedattention", identity=False):
    """The reference on the host CPU, fp32, all host threads: the UNMODIFIED reference staged in oracle/_ref
    (oracle/build_ref.py) driven through its own Model.train_paired() / train_cycle() -- kind "reference" -- or, when it
    has not been staged, the oracle's restatement of the same loop -- kind "port". Returns (tiles/s, seconds, cores,
    kind)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cycle = model in CYCLE_MODELS
    from oracle import ref_runner
    if ref_runner.available():
        m = ref_runner.make_model(model, add_identity_loss=identity)
        if warmup:
            ref_runner.run_epoch(m, _synthetic_host_batches(0, warmup, batch), cycle)
        data = _synthetic_host_batches(warmup, steps, batch)
        t0 = time.perf_counter()
        ref_runner.run_epoch(m, data, cycle)
        dt = time.perf_counter() - t0
        return batch * steps / dt, dt, cores, "reference"
    from oracle import gan_oracle as O
    nets = O.init_model(model, "all", seed=47)
    tr = O.CycleTrainer(nets, model, add_identity_loss=identity) if cycle else O.PairedTrainer(nets, model)
    for s in range(warmup):
        tr.step(*O.synthetic_batch(s, batch, CHANNELS, TILE))
    data = [O.synthetic_batch(warmup + s, batch, CHANNELS, TILE) for s in range(steps)]
    t0 = time.perf_counter()
    for x, y in data:
        tr.step(x, y)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt, cores, "port"


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the step on this box's host cores, same `config`,
    metric and unit as the native arm. A step is a bounded sample of the batch-16 workload: the per-step batch is the
    largest of 16 / 8 / 4 / 2 / 1 tiles for which warm-up + timed steps fit in about three minutes (probed with one
    batch-1 step); throughput is tiles/s either way."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    probe, _, _, _ = cpu_reference_tiles_per_s(1, 1, 1, args.model, args.identity)
    batch = BATCH_PER_GPU
    while batch > 1 and (steps + warmup) * batch / probe > 180.0:
        batch //= 2
    tps, dt, cores, kind = cpu_reference_tiles_per_s(steps, warmup, batch, args.model, args.identity)
    source = ("the unmodified reference (oracle/_ref, staged by oracle/build_ref.py) through its own Model."
              f"{'train_cycle' if args.model in CYCLE_MODELS else 'train_paired'}()" if kind == "reference" else
              "oracle port (oracle/gan_oracle.py) of the reference's training loop")
    sample = (f"{source}, fp32, {cores} host threads, {batch} of the {BATCH_PER_GPU} tiles of a batch per step, "
              f"{steps} timed steps after {warmup} warm-up")
    line = {"impl": "reference", "metric": f"{PRETTY[args.model]} train tiles/s", "value": tps, "unit": "tiles/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1000 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.model, args.identity, BATCH_PER_GPU, args.gpus),
            "cpu_baseline": {"value": tps, "unit": "tiles/s", "cores": cores, "kind": kind, "sample": sample},
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
        # an NCCL_DEBUG set by the caller is honoured (stdout is already diverted to stderr, see main())
        os.environ.setdefault("NCCL_DEBUG", "WARN")
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

    cycle = args.model in CYCLE_MODELS
    model = M.Model(model=PRETTY[args.model], topography="all", num_epochs=1, resize=512, crop=4, seed=47,
                    log_interval=1, add_identity_loss=args.identity)
    tr = model._ensure_native_cycle() if cycle else model._ensure_native_paired()
    train_api = model.train_cycle if cycle else model.train_paired
    gflop_per_tile = GFLOP_PER_TILE[(args.model, args.identity)]

    # ---- (1) device-resident inputs: `value`
    loader = SyntheticLoader(steps=4, batch=B, channels=CHANNELS, size=TILE, rank=rank, world_size=world, pin=False)
    resident = [(x.to(dev), y.to(dev)) for x, y, _ in loader]
    for s in range(warmup):
        tr.step(*resident[s % len(resident)])
    barrier()
    ops.LAUNCHES = 0
    peer_reds = [r for r in (getattr(tr, "g_reducer", None), getattr(tr, "d_reducer", None)) if hasattr(r, "waited_ms")]
    waited0 = sum(r.waited_ms() for r in peer_reds)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for s in range(steps):
            tr.step(*resident[s % len(resident)])
        e1.record()
        barrier()
    launches = ops.LAUNCHES
    waited_ms_per_step = (sum(r.waited_ms() for r in peer_reds) - waited0) / steps if peer_reds else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = B * world * steps / (ms / 1000.0)
    last_losses = tr.losses()

    # ---- (2) end to end through the public API: Model.train_paired with pinned HOST batches; every step copies its
    # inputs host->device and reads the four losses back (log_interval=1)
    host = list(SyntheticLoader(steps=steps, batch=B, channels=CHANNELS, size=TILE, rank=rank, world_size=world))
    model.train_loader = host[:min(3, steps)]
    model.num_epochs = model.starting_epoch = 1
    train_api()  # warm-up epoch of the API path
    model.train_loader = host
    model.starting_epoch = model.num_epochs = 2
    barrier()
    e0.record()
    t0 = time.perf_counter()
    train_api()
    e1.record()
    barrier()
    wall_ms = 1000 * (time.perf_counter() - t0)
    e2e_event_ms = e0.elapsed_time(e1)
    e2e_ms = max_over_ranks(max(e2e_event_ms, wall_ms))
    e2e = B * world * steps / (e2e_ms / 1000.0)
    # diagnostic: pinned host->device bandwidth of this box for one batch (the e2e path hides the copy of batch i+1
    # behind step i; that only works while a batch copies faster than a step runs)
    hx, hy = host[0][0], host[0][1]
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        hx.to(dev, non_blocking=True)
        hy.to(dev, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_ms_per_batch = c0.elapsed_time(c1) / 3
    h2d_gbs = (hx.numel() + hy.numel()) * 4 / (h2d_ms_per_batch * 1e6)

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
                    "traffic": ncu_traffic_bytes(),
                    "traffic_source": "committed ncu --set full capture of this kernel on this shape "
                                      "(profiles/r02/ncu_kernels.csv), not measured in this run",
                    "peak_source": peak_src + ", sustained bf16", "launches_timed": len(res_ms),
                    "avg_ms": avg}
    conv_ms = sum(sum(v) for k, v in prof.items() if k.split()[0] in ("fprop", "dgrad", "wgrad")) / 2
    all_ms = sum(sum(v) for v in prof.values()) / 2
    step_tflops = gflop_per_tile * B * world / (ms / steps)

    extra = {}
    if rank == 0 and world == 1 and args.model == "pairedattention" and not args.no_unet:
        extra["unet"] = unet_block(dev)
    if rank == 0:
        cpu_tps = None
        if world == 1:  # bounded sample: one warm-up + two timed batch-16 steps (paired: ~10-20 s of host work)
            cpu_b = BATCH_PER_GPU if not cycle else 4
            cpu_tps, cpu_dt, cores, cpu_kind = cpu_reference_tiles_per_s(2, 1, cpu_b, args.model, args.identity)
        line = {"metric": f"{PRETTY[args.model]} train tiles/s", "value": value, "unit": "tiles/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args.model, args.identity, B, world),
                "clocks": clocks.summary(),
                "e2e": {"value": e2e, "unit": "tiles/s", "h2d_bytes_per_step": B * (CHANNELS + 3) * TILE * TILE * 4,
                        "d2h_bytes_per_step": 4 * len(last_losses), "ms_per_step": e2e_ms / steps,
                        "ms_per_step_device_events": e2e_event_ms / steps, "ms_per_step_host_wall": wall_ms / steps,
                        "h2d_pinned_ms_per_batch": h2d_ms_per_batch, "h2d_pinned_GBps": h2d_gbs,
                        "host_batches_pinned": bool(hx.is_pinned()),
                        "api": f"models.model.Model.{'train_cycle' if cycle else 'train_paired'}() with pinned host "
                               "batches, log_interval=1"},
                "gpu_launches": launches,
                "roofline": roofline,
                "step_tflops": step_tflops, "step_frac_of_peak": step_tflops / (peak_tf * world),
                "conv_share_of_step": conv_ms / all_ms if all_ms else None,
                "kernel_ms_per_step": {k: sum(v) / 2 for k, v in sorted(prof.items(), key=lambda kv: -sum(kv[1]))[:12]},
                "losses_last_step": last_losses, "extra": extra}
        red = getattr(getattr(model, "_native", None), "g_reducer", None)
        if red is not None:  # how the gradients of the ranks are summed (fpgan/peer.py or NCCL)
            peer = type(red).__name__ == "PeerReducer"
            line["gradient_exchange"] = {
                "kind": "copy-engine all-gather over peer memory + rank-ordered sum inside Adam" if peer
                        else "NCCL all-reduce",
                "buckets": len(red.bounds), "bytes_pushed_per_rank_per_step":
                    sum(4 * r.count * (world - 1) for r in (model._native.g_reducer, model._native.d_reducer))
                    if peer else None,
                # rank 0's stream time inside the flag waits of the timed (replayed) steps, device clock: the part of
                # the exchange the backward pass did not hide + the skew between the ranks
                "wait_ms_per_step_rank0": waited_ms_per_step,
                "adam_ms_per_step_eager": sum(prof.get("adam_step", [])) / 2}
        if cpu_tps is not None:
            line["cpu_baseline"] = {"value": cpu_tps, "unit": "tiles/s", "cores": cores, "kind": cpu_kind,
                                    "sample": ("the unmodified reference (oracle/_ref) through its own Model training "
                                               "loop" if cpu_kind == "reference" else "oracle port of the training "
                                               "loop") + f", fp32, 2 timed batch-{cpu_b} steps (256x256) after 1 "
                                              "warm-up"}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    model.close()  # captured NCCL work must be gone before the communicator is torn down
    if world > 1:
        dist.destroy_process_group()


def unet_block(dev):
    """BASELINE configs[4] (reported beside the headline): segmentation U-Net inference +
    bit-exact flood masks + confusion counts on generated vs ground-truth 256x256 tiles, batch 64 -- tile pairs/s with
    CUDA events -- and, on a bounded sample (8 tile pairs), the agreement of the masks and confusion counts with the fp32
    CPU oracle (BatchNorm uses batch statistics, so both sides see the same batch). The U-Net is randomly initialised
    (no checkpoint offline): its logits sit near the threshold, the disagreement is the band bf16 cannot decide."""
    from models import model as M
    from models import model_architectures as A
    torch.manual_seed(47)
    net = A.UNet().apply(M.Model.initialise_weights)  # as Model.load_segmentation_model without a checkpoint
    p = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(dev)
    g = torch.Generator().manual_seed(2000)
    gen = (torch.rand(64, 3, TILE, TILE, generator=g) * 2 - 1).to(dev)
    truth = (torch.rand(64, 3, TILE, TILE, generator=g) * 2 - 1).to(dev)
    for _ in range(2):
        A.flood_masks_and_counts(net, gen, truth)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters):
        A.flood_masks_and_counts(net, gen, truth)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    n = 8
    mo, mt, counts = A.flood_masks_and_counts(net, gen[:n].contiguous(), truth[:n].contiguous())
    from oracle import gan_oracle as O  # the checker: fp32 CPU restatement of the reference's U-Net + threshold
    with torch.no_grad():
        ro, rt = O.segmentation_masks(p, gen[:n].cpu(), truth[:n].cpu())
    ref_counts = O.confusion_counts(ro.flatten(), rt.flatten())
    return {"workload": "segmentation U-Net inference on generated + ground-truth 256x256 tiles, batch 64, bit-exact "
                        "(sigmoid > 0.5) masks, TP/FP/TN/FN counts",
            "tile_pairs_per_s": 64 / (ms / 1000.0), "ms_per_batch": ms,
            "tflops_algorithmic": 2 * 96.33 * 64 / ms,
            "oracle_sample_pairs": n,
            "mask_disagreement_rate_generated": (mo.cpu() != ro).float().mean().item(),
            "mask_disagreement_rate_truth": (mt.cpu() != rt).float().mean().item(),
            "confusion_counts_native": counts.tolist(), "confusion_counts_oracle": ref_counts}


def run_dp_check(args):
    """Data-parallel parity on hardware: W ranks x B/W tiles against ONE process x B tiles, same
    seeds, three steps (two eager, one captured + replayed). Every rank computes the single-process run itself (no
    collective), then the ranks run the sharded steps together. Reported: per-step max relative difference of the
    (rank-averaged) losses and the relative RMS difference of all generator / discriminator weights at the end."""
    import torch.distributed as dist
    from fpgan import trainer as T
    from models import model as M
    from models.data import SyntheticLoader

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    assert B % world == 0, "--batch must divide by the number of ranks"
    cycle = args.model in CYCLE_MODELS
    size = args.check_size

    def build(world_size):
        m = M.Model(model=PRETTY[args.model], topography="all", num_epochs=1, resize=512, crop=4, seed=47,
                    add_identity_loss=args.identity)
        if cycle:
            return m, T.CycleTrainer(m.pre_to_post_generator, m.post_to_pre_generator, m.pre_discriminator,
                                     m.post_discriminator, add_identity_loss=args.identity, world_size=world_size)
        return m, T.PairedTrainer(m.generator, m.discriminator, world_size=world_size)

    def snapshot(tr, out, w):
        img = (out[0] if isinstance(out, tuple) else out).clone()
        return img, tr.gp.grads.flat.clone() / w, tr.dp.grads.flat.clone() / w

    def run(tr, batch, r, w):
        """three steps; returns per-step losses, the first step's generated images and its parameter gradients (the
        all-reduced SUM over ranks, scaled here by 1/w as Adam does)"""
        hist, first = [], None
        for x, y, _ in SyntheticLoader(steps=3, batch=batch, channels=CHANNELS, size=size, rank=r, world_size=w,
                                       pin=False):
            out = tr.step(x.to(dev), y.to(dev))
            hist.append(tr.losses())
            if first is None:
                first = snapshot(tr, out, w)
        return hist, first

    def run_accumulated(tr, w):
        """the same three steps as ONE process stepping over the w shards (gradient accumulation in shard order)"""
        hist, first = [], None
        loaders = [iter(SyntheticLoader(steps=3, batch=B // w, channels=CHANNELS, size=size, rank=r, world_size=w,
                                        pin=False)) for r in range(w)]
        for _ in range(3):
            shards = [next(it) for it in loaders]
            outs = tr.step_accumulated([(x.to(dev), y.to(dev)) for x, y, _ in shards])
            hist.append(tr.losses())
            if first is None:
                first = snapshot(tr, outs[rank], 1)  # gradients already summed over the shards; Adam scales by 1/w
                first = (first[0], first[1] / w, first[2] / w)
        return hist, first

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-30)).item()

    def compare(ref_hist, ref_first, ref_w, hist, first, w_flat, rows):
        diffs = [max(abs(a[k] - c[k]) / (abs(a[k]) + 1e-12) for k in a) for a, c in zip(ref_hist, hist)]
        img = ref_first[0] if rows is None else ref_first[0][rows]
        return {"loss_max_rel_diff_per_step": diffs, "step0_generated_rel_rms_diff": rel(first[0], img),
                "step0_generator_grad_rel_rms_diff": rel(first[1], ref_first[1]),
                "step0_discriminator_grad_rel_rms_diff": rel(first[2], ref_first[2]),
                "weights_rel_rms_diff_after_3_steps": rel(w_flat, ref_w)}

    b = B // world
    # (a) ONE process, the same shards one after the other (identical kernels per shard): must agree to fp32
    #     summation-order noise -- this isolates the data-parallel machinery (sharding, all-reduce, 1/W, loss averaging)
    m0, t0 = build(1)
    acc_hist, acc_first = run_accumulated(t0, world)
    w_acc = torch.cat([t0.gp.flat, t0.dp.flat]).clone()
    del t0, m0
    # (b) ONE process, the whole global batch at once (the reference's formulation): kernel plans -- tile shapes, split-K
    #     factors, partial-statistics rows -- depend on the per-process batch, so fp32 sums are taken in another order; a
    #     1e-7 difference flips the bf16 / fp16 rounding of a few stored activations and the rounding noise of the two
    #     runs decorrelates layer by layer (informational: both are equally far from the fp32 oracle)
    m1, t1 = build(1)
    single, single_first = run(t1, B, 0, 1)
    w_single = torch.cat([t1.gp.flat, t1.dp.flat]).clone()
    del t1, m1
    os.environ["FPG_PEER_KEEP_SUM"] = "1"  # leave the rank-ordered gradient SUM in grads.flat (compared below)
    mw, tw = build(world)
    sharded, sh_first = run(tw, b, rank, world)
    w_sharded = torch.cat([tw.gp.flat, tw.dp.flat])
    vs_acc = compare(acc_hist, acc_first, w_acc, sharded, sh_first, w_sharded, None)
    vs_single = compare(single, single_first, w_single, sharded, sh_first, w_sharded, slice(rank * b, (rank + 1) * b))
    tol = {"step0_loss": 1e-5, "step0_generated": 1e-6, "step0_grads": 1e-4, "later_losses": 2e-3, "weights": 2e-3}
    ok = (vs_acc["loss_max_rel_diff_per_step"][0] <= tol["step0_loss"] and
          max(vs_acc["loss_max_rel_diff_per_step"]) <= tol["later_losses"] and
          vs_acc["step0_generated_rel_rms_diff"] <= tol["step0_generated"] and
          vs_acc["step0_generator_grad_rel_rms_diff"] <= tol["step0_grads"] and
          vs_acc["step0_discriminator_grad_rel_rms_diff"] <= tol["step0_grads"] and
          vs_acc["weights_rel_rms_diff_after_3_steps"] <= tol["weights"])
    if world > 1:  # every rank must agree
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() > 0)
    rec = {"vs_one_process_accumulating_the_same_shards": vs_acc, "vs_one_process_whole_batch": vs_single}
    if rank == 0:
        print(json.dumps({"check": "data-parallel parity", "model": args.model, "identity": args.identity,
                          "world": world, "global_batch": B, "tile": size, "steps": 3, **rec, "tolerance": tol,
                          "ok": bool(ok), "losses_single": single[-1], "losses_sharded": sharded[-1]}),
              file=_RESULT_OUT, flush=True)
    tw.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(3)


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
    ap.add_argument("--model", default="pairedattention", choices=sorted(PRETTY),
                    help="pairedattention (BASELINE configs[1]/[2], default), the cycle models of configs[3], or pix2pix "
                         "(the model of configs[0], fused step)")
    ap.add_argument("--identity", action="store_true", help="cycle models: add the identity loss (model.py:700-702)")
    ap.add_argument("--check", action="store_true", help="data-parallel parity check instead of a benchmark")
    ap.add_argument("--no_unet", action="store_true", help="skip the extra.unet block (configs[4]) of the 1-GPU line")
    ap.add_argument("--check_size", type=int, default=TILE)
    args = ap.parse_args()
    args.model = args.model.lower()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.check:
        run_dp_check(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
