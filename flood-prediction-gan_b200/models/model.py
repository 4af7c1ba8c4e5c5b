"""`Model`: the training / evaluation orchestrator with the reference's constructor arguments, attributes, loss
bookkeeping and checkpoint layout (reference: models/model.py:26-160, 296-361, 598-758), driving the native
sm_100a executors. `train.py` / `evaluate.py` pass `Model(**vars(args))` exactly as the reference does.

Differences by design (documented in DESIGN.md): the paired step runs fused through fpgan.trainer (no autograd
graph, losses stay on the device and are read back once per `log_interval` steps instead of 4x per step), the
batch can be larger than 1 and sharded over GPUs (torch.distributed / NCCL), and plotting / torchmetrics-based
image-quality metrics are out of scope.
"""
import itertools
import os
import random
import time
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist
from torch import nn
from torch.optim import lr_scheduler

from fpgan import trainer as native_trainer
from models import data, metrics, model_architectures

MODEL_TABLE = {
    # name: (generator, discriminator, cycle training?, attention generator?)
    "pix2pix": ("Pix2PixGenerator", "Pix2PixDiscriminator", False, False),
    "pairedattention": ("PairedAttentionGenerator", "PairedAttentionDiscriminator", False, True),
    "cyclegan": ("CycleGANGenerator", "CycleGANDiscriminator", True, False),
    "attentiongan": ("AttentionGANGenerator", "AttentionGANDiscriminator", True, True),
}
PRETTY = {"pix2pix": "Pix2Pix", "cyclegan": "CycleGAN", "attentiongan": "AttentionGAN",
          "pairedattention": "PairedAttention"}
TOPOGRAPHY_CHANNELS = {"all": 9, "map": 6, "dem": 4, "flow": 4, "river": 4, None: 3}
NOT_IMPLEMENTED_MSG = "Model must be one of: Pix2Pix, CycleGAN, AttentionGAN or PairedAttention"


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: this implementation runs only on sm_100a GPUs (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class Model:
    def __init__(self, model="Pix2Pix", dataset_subset="all", dataset_dem="best", data_path=None, num_epochs=1,
                 topography="all", resize=256, crop=None, save_model_interval=0, save_images_interval=0,
                 verbose=False, load_pretrained_model=False, pretrained_model_path=None, add_identity_loss=False,
                 training_model=True, seed=47, log_interval=50, batch_size=1):
        self.device = _device()
        if verbose:
            print(f"\nSetting up the {self.prettify_model_name(model)} model...")
        saved = None
        if load_pretrained_model:
            # reference checkpoints hold numpy scalars in all_losses -> weights_only=False (trusted local file)
            saved = torch.load(pretrained_model_path, map_location=self.device, weights_only=False)
            model, num_epochs = saved["model"], saved["num_epochs"]
            topography, add_identity_loss = saved["topography"], saved["add_identity_loss"]
        self.model = model.lower()
        self.num_epochs, self.topography, self.add_identity_loss = num_epochs, topography, add_identity_loss
        self.verbose, self.save_model_interval, self.save_images_interval = verbose, save_model_interval, save_images_interval
        self.load_pretrained_model, self.data_path = load_pretrained_model, data_path
        self.dataset_subset, self.dataset_dem, self.resize, self.crop = dataset_subset, dataset_dem, resize, crop
        self.training_model, self.seed, self.log_interval = training_model, seed, log_interval
        if self.model not in MODEL_TABLE:
            raise NotImplementedError(NOT_IMPLEMENTED_MSG)
        gen_name, dis_name, self.model_is_cycle, self.model_is_attention = MODEL_TABLE[self.model]
        gen_cls = getattr(model_architectures, gen_name, None)
        dis_cls = getattr(model_architectures, dis_name, None)
        if gen_cls is None or dis_cls is None:
            raise NotImplementedError(f"{PRETTY[self.model]} is not available in this build yet")

        n_in = TOPOGRAPHY_CHANNELS[self.topography]
        torch.manual_seed(self.seed)  # one seed, then G(s) and D(s) in the reference's order (model.py:80-104)

        def make(cls):
            return cls(input_channels=n_in).apply(self.initialise_weights).to(self.device)

        if self.model_is_cycle:
            self.pre_to_post_generator, self.post_to_pre_generator = make(gen_cls), make(gen_cls)
            if training_model:
                self.pre_discriminator, self.post_discriminator = make(dis_cls), make(dis_cls)
        else:
            self.generator = make(gen_cls)
            if training_model:
                self.discriminator = make(dis_cls)

        if training_model:
            self.loss_func = nn.MSELoss()
            if self.model_is_cycle:
                self.cycle_loss, self.identity_loss = nn.L1Loss(), nn.L1Loss()
                g_params = itertools.chain(self.pre_to_post_generator.parameters(),
                                           self.post_to_pre_generator.parameters())
                d_params = itertools.chain(self.post_discriminator.parameters(), self.pre_discriminator.parameters())
            else:
                self.l1_loss = nn.L1Loss()
                g_params, d_params = self.generator.parameters(), self.discriminator.parameters()
            # torch optimisers are kept for their state_dict layout (checkpoints); the fused paired step updates
            # the same parameters with the native Adam kernel and mirrors its moments into these on save
            self.optimizer_discriminator = torch.optim.Adam(d_params, lr=0.0002, betas=(0.5, 0.999))
            self.optimizer_generator = torch.optim.Adam(g_params, lr=0.0002, betas=(0.5, 0.999))
            self.scheduler_generator = lr_scheduler.LambdaLR(self.optimizer_generator, lr_lambda=self.lambda_rule)
            self.scheduler_discriminator = lr_scheduler.LambdaLR(self.optimizer_discriminator,
                                                                 lr_lambda=self.lambda_rule)

        if saved is not None:
            self.starting_epoch, self.all_losses = saved["starting_epoch"], saved["all_losses"]
            if training_model:
                self.optimizer_discriminator.load_state_dict(saved["optimizer_discriminator"])
                self.optimizer_generator.load_state_dict(saved["optimizer_generator"])
                self.scheduler_discriminator.load_state_dict(saved["scheduler_discriminator"])
                self.scheduler_generator.load_state_dict(saved["scheduler_generator"])
            for key in self._network_names():
                getattr(self, key).load_state_dict(saved[key])
        else:
            self.starting_epoch = 1
            self.all_losses = self.initialise_loss_storage(overall=True)
        self.current_epoch = self.starting_epoch
        self._native = None

        # device-resident loaders (models/data.py); batch_size is an extension, the reference's loaders use 1
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        self.train_loader, self.val_loader, self.test_loader = data.create_flood_dataset(
            self.dataset_subset, self.dataset_dem, self.data_path, self.topography, self.resize, self.crop,
            batch_size=batch_size, rank=rank, world_size=world)
        if self.verbose and self.training_model:
            self.print_training_setup()

    # ------------------------------------------------------------------------------------------ helpers
    def _network_names(self):
        if self.model_is_cycle:
            names = ["pre_to_post_generator", "post_to_pre_generator"]
            if self.training_model:
                names += ["pre_discriminator", "post_discriminator"]
        else:
            names = ["generator"] + (["discriminator"] if self.training_model else [])
        return names

    @staticmethod
    def initialise_weights(m):
        """N(0, 0.02) conv / linear weights with zero bias, N(1, 0.02) BatchNorm weights (reference :162-173)."""
        name = m.__class__.__name__
        if hasattr(m, "weight") and ("Conv" in name or "Linear" in name):
            nn.init.normal_(m.weight.data, 0.0, 0.02)
            if getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif "BatchNorm2d" in name:
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0.0)

    def lambda_rule(self, epoch):
        """Constant lr for the first half of the epochs, linear decay over the second half (reference :175-181)."""
        half = self.num_epochs / 2
        return 1.0 - max(0, epoch + 1 - half) / float(half + 1)

    def initialise_loss_storage(self, overall):
        pre = "all_" if overall else ""
        if self.model_is_cycle:
            keys = ["losses_generator_post", "losses_generator_pre", "losses_pre_to_post_cycle",
                    "losses_post_to_pre_cycle", "losses_discriminator_pre_real", "losses_discriminator_post_real",
                    "losses_discriminator_pre_synthetic", "losses_discriminator_post_synthetic"]
            if self.add_identity_loss:
                keys += ["losses_identity_post", "losses_identity_pre"]
        else:
            keys = ["losses_discriminator_real", "losses_discriminator_synthetic", "losses_generator_synthetic",
                    "l1_losses_generator_synthetic"]
        return {pre + k: [] for k in keys}

    def prettify_model_name(self, model_name=None):
        return PRETTY[model_name.lower()] if model_name else PRETTY[self.model]

    def create_path(self, save_type, info=""):
        ext = {"image": ".png", "figure": ".png", "model": ".pth.tar", "metric": ".csv"}[save_type]
        stamp = str(datetime.now())[:-7].replace(" ", "-").replace(":", "-")
        idt = f"identity{self.add_identity_loss}" if self.model_is_cycle else ""
        epoch = self.current_epoch if self.training_model else self.current_epoch - 1
        path = (f"{self.data_path}/{save_type}s/{self.prettify_model_name()}_{info}_epoch{epoch}_"
                f"{self.topography}Topography_{idt}_{self.dataset_subset}Data_{self.dataset_dem}DEM_"
                f"resize{self.resize}_crop{self.crop}_date{stamp}{ext}")
        return path.replace("__", "_")

    def print_training_setup(self):
        print(f"\n{'Continuing' if self.load_pretrained_model else 'Beginning'} training {self.prettify_model_name()}:")
        print(f"{self.num_epochs} epochs\nStarting from epoch {self.starting_epoch}")
        print(f"{self.topography.title() if self.topography else 'No'} topographical factors will be input to the model")
        if self.model_is_cycle and self.add_identity_loss:
            print("Using identity mapping loss")
        print(f"Dataset: {len(self.train_loader)} images from '{self.dataset_subset}' with '{self.dataset_dem}' DEM")
        print(f"Data resized to {self.resize} pixels with {self.crop} crops, scaled to [-1, 1]")
        print(f"Model saved every {self.save_model_interval} epochs")
        print(f"Sample generator output images saved every {self.save_images_interval} epochs\n")

    def get_buffer_image(self, image, images_buffer):
        """50-entry history buffer of generated images (reference :275-294); kept on the device."""
        image = image.detach()
        if len(images_buffer) < 50:
            images_buffer.append(image.clone())
            return image
        if random.uniform(0, 1) > 0.5:
            index = random.randint(0, 49)
            old = images_buffer[index].clone()
            images_buffer[index] = image.clone()
            return old
        return image

    def print_losses(self):
        last = {k: v[-1] for k, v in self.all_losses.items()}
        print("| " + " | ".join(f"{k[4:]} = {v:.2f}" for k, v in last.items()))

    # ------------------------------------------------------------------------------------------ checkpoints
    def _sync_optimizer_state_for_save(self):
        """Mirror the native flat Adam moments into the torch optimisers so that their state_dict() has the
        reference layout ({'state': {i: {step, exp_avg, exp_avg_sq}}, 'param_groups': [...]})."""
        if self._native is None:
            return
        for fp, opt in ((self._native.gp, self.optimizer_generator), (self._native.dp, self.optimizer_discriminator)):
            for name, p in fp.named:
                off, k = fp.offsets[name]
                opt.state[p] = {"step": torch.tensor(float(fp.steps)),
                                "exp_avg": fp.m[off:off + k].view(p.shape).clone(),
                                "exp_avg_sq": fp.v[off:off + k].view(p.shape).clone()}

    def save_results(self, epoch, losses, epoch_start_time):
        self.current_epoch = epoch
        for key in self.all_losses:
            self.all_losses[key].append(np.mean(losses[key[4:]]))
        if self.verbose:
            print(f"Epoch {epoch} ({time.time() - epoch_start_time:.2f} seconds) ", end="")
            self.print_losses()
        if self.save_model_interval != 0 and epoch % self.save_model_interval == 0 and self._rank() == 0:
            # data-parallel replicas hold identical weights and optimiser state: rank 0 writes the checkpoint
            self._sync_optimizer_state_for_save()
            saved = {"model": self.model, "starting_epoch": epoch + 1, "num_epochs": self.num_epochs,
                     "topography": self.topography,
                     "optimizer_generator": self.optimizer_generator.state_dict(),
                     "optimizer_discriminator": self.optimizer_discriminator.state_dict(),
                     "scheduler_generator": self.scheduler_generator.state_dict(),
                     "scheduler_discriminator": self.scheduler_discriminator.state_dict(),
                     "all_losses": self.all_losses, "add_identity_loss": self.add_identity_loss}
            for key in self._network_names():
                saved[key] = getattr(self, key).state_dict()
            path = self.create_path(save_type="model")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            print(f"Saving {self.prettify_model_name()} model to {path}")
            torch.save(saved, path)


    # ------------------------------------------------------------------------------------------ evaluation
    @staticmethod
    def binary_metrics_from_counts(tp, fp, tn, fn):
        """torchmetrics Binary{Accuracy,F1Score,Precision,Recall} and MeanSquaredError of two {0,1} masks restated from
        the integer confusion counts (reference call sites model.py:372-378, 409-418; torchmetrics divides safely:
        0 when a denominator is 0). The *_No_Flood metrics are the same formulas on the inverted masks abs(mask - 1)
        (model.py:415-416), i.e. with tp<->tn and fp<->fn exchanged."""
        def div(a, b):
            return float(a) / float(b) if b else 0.0
        n = tp + fp + tn + fn
        return {"MSE": div(fp + fn, n), "Accuracy": div(tp + tn, n),
                "F1_Flood": div(2 * tp, 2 * tp + fp + fn), "Precision_Flood": div(tp, tp + fp),
                "Recall_Flood": div(tp, tp + fn),
                "F1_No_Flood": div(2 * tn, 2 * tn + fn + fp), "Precision_No_Flood": div(tn, tn + fn),
                "Recall_No_Flood": div(tn, tn + fp)}

    @staticmethod
    def metrics_csv_text(results):
        """the text the reference's `metrics_df.to_csv()` writes (model.py:419-421): unnamed index column, one row
        labelled 1, NaN as an empty field"""
        cells = ["" if v != v else str(v) for v in results.values()]
        return "," + ",".join(results.keys()) + "\n1," + ",".join(cells) + "\n"

    def load_segmentation_model(self, seg_model_path=None):
        """SegmentationModel(train=False).model (segmentation_model.py:55-61): U-Net initialised like the reference
        and, if a checkpoint is given, loaded from its "model" entry. Never switched to eval mode (the reference
        does not)."""
        net = model_architectures.UNet().apply(self.initialise_weights).to(self.device)
        if seg_model_path:
            saved = torch.load(seg_model_path, map_location=self.device, weights_only=False)
            net.load_state_dict(saved["model"])
        return net

    def calculate_metrics(self, use_test_data=False, seg_model_path=None, loader=None):
        """Flood-segmentation metrics of calculate_metrics (reference model.py:363-422): generator inference,
        segmentation of the generated and the ground-truth tile, bit-exact (sigmoid > 0.5) masks, confusion counts
        accumulated over the whole split on the device; PSNR / SSIM / MS-SSIM per batch from models/metrics.py (the
        published torchmetrics algorithm; that dependency is absent, parity unpinned), averaged over batches as the
        reference does with np.mean; LPIPS is NaN. Returns the metrics dict (and writes the reference's CSV when
        data_path is set)."""
        seg_model = self.load_segmentation_model(seg_model_path)
        generator = self.pre_to_post_generator if self.model_is_cycle else self.generator
        if loader is None:
            loader = self.test_loader if use_test_data else self.val_loader
        totals = torch.zeros(4, dtype=torch.int64, device=self.device)
        times = []
        quality = {"PSNR": metrics.PeakSignalNoiseRatio(data_range=(0, 1)),                       # :367-369
                   "SSIM": metrics.StructuralSimilarityIndexMeasure(data_range=(0, 1)),
                   "MS-SSIM": metrics.MultiScaleStructuralSimilarityIndexMeasure(data_range=(0, 1))}
        per_batch = {k: [] for k in quality}
        n_in = TOPOGRAPHY_CHANNELS[self.topography]
        for input_stack, ground_truth, _ in loader:
            x = input_stack[:, :n_in].to(self.device).float().contiguous()
            truth = ground_truth.to(self.device).float()
            torch.cuda.synchronize()
            t0 = time.time()
            torch.manual_seed(47)
            with torch.no_grad():
                generated = generator(x)
            torch.cuda.synchronize()
            times.append(time.time() - t0)
            _, _, counts = model_architectures.flood_masks_and_counts(seg_model, generated, truth)
            totals += counts
            g01 = torch.clamp((generated + 1) * 0.5, min=0, max=1)                                # :397-398
            t01 = torch.clamp((truth + 1) * 0.5, min=0, max=1)
            for name, metric in quality.items():                                                    # :404-406
                if name == "MS-SSIM" and not metrics.ms_ssim_size_ok(*g01.shape[-2:]):
                    continue  # torchmetrics raises for tiles this small; the column stays NaN
                per_batch[name].append(metric(g01, t01))
                metric.reset()
        tp, fp, tn, fn = (int(v) for v in totals.tolist())
        results = {k: (float(torch.stack(v).mean().item()) if v else float("nan")) for k, v in per_batch.items()}
        results["LPIPS"] = float("nan")  # needs pretrained AlexNet weights (download): not provided
        results.update(self.binary_metrics_from_counts(tp, fp, tn, fn))
        results["Inference"] = float(np.mean(times)) if times else float("nan")
        if self.verbose:
            print(results)
        if self.data_path and os.path.isdir(str(self.data_path)):
            path = self.create_path("metric")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(self.metrics_csv_text(results))
        return results

    # ------------------------------------------------------------------------------------------ training
    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
        return 1

    def _rank(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
        return 0

    def _adopt_optimizer_state(self):
        """resume: adopt Adam moments stored in the torch optimisers (reference checkpoint layout)"""
        for fp, opt in ((self._native.gp, self.optimizer_generator), (self._native.dp, self.optimizer_discriminator)):
            for name, p in fp.named:
                st = opt.state.get(p)
                if st:
                    off, k = fp.offsets[name]
                    fp.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    fp.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    fp.steps = int(st["step"])

    def _ensure_native_paired(self):
        if self._native is None:
            cls = native_trainer.Pix2PixTrainer if self.model == "pix2pix" else native_trainer.PairedTrainer
            self._native = cls(self.generator, self.discriminator, world_size=self._world())
            self._adopt_optimizer_state()
        return self._native

    def _ensure_native_cycle(self):
        if self._native is None:
            self._native = native_trainer.CycleTrainer(
                self.pre_to_post_generator, self.post_to_pre_generator, self.pre_discriminator,
                self.post_discriminator, add_identity_loss=self.add_identity_loss, world_size=self._world())
            self._adopt_optimizer_state()
        return self._native

    def close(self):
        """Release the captured CUDA graphs of the fused steps. With data parallelism they hold NCCL work: call this
        (then synchronise and barrier) before torch.distributed.destroy_process_group(), which hangs otherwise."""
        if self._native is not None:
            self._native.close()  # captured steps, then the peer-memory mappings of the gradient exchange (collective)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if self._world() > 1:
            dist.barrier()

    def _allreduce_module_grads(self, params):
        """data parallelism of the module (autograd) path: sum the parameter gradients over ranks, average"""
        world = self._world()
        if world == 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def train_paired(self):
        """Paired training (reference :598-658): the fused native step (fpgan.trainer.PairedTrainer for PairedAttention,
        Pix2PixTrainer for Pix2Pix -- BatchNorm with batch statistics, dropout). FPG_PAIRED_MODULES=1 selects the
        reference loop over the drop-in modules' autograd path instead (cross-check)."""
        if os.environ.get("FPG_PAIRED_MODULES", "0") == "1":
            return self._train_paired_modules()
        tr = self._ensure_native_paired()
        for epoch in range(self.starting_epoch, self.num_epochs + 1):
            t0 = time.time()
            losses = self.initialise_loss_storage(overall=False)
            self.discriminator.train()
            self.generator.train()
            torch.manual_seed(epoch)
            lr_g = self.optimizer_generator.param_groups[0]["lr"]
            lr_d = self.optimizer_discriminator.param_groups[0]["lr"]
            pending = []
            for x, y in self._prefetch_to_device(self.train_loader):
                tr.step(x, y, lr_g=lr_g, lr_d=lr_d)
                pending.append(tr.loss_buf.clone())  # stays on the device
                # losses are read back one step late: the device->host read of step i's losses is issued after step
                # i+1 has been enqueued, so the GPU never waits for the host between steps
                if len(pending) > self.log_interval:
                    self._flush_losses(pending, losses, keep_last=1)
            self._flush_losses(pending, losses)
            self.scheduler_discriminator.step()
            self.scheduler_generator.step()
            self.save_results(epoch=epoch, losses=losses, epoch_start_time=t0)

    def _prefetch_to_device(self, loader):
        """Yields device batches with the NEXT batch's host->device copies already in flight on a side stream (the
        50 MB of a batch-16 step take 1-4 ms over PCIe: hidden behind the previous step instead of serialised with it).
        Copies are asynchronous only from pinned host memory, which the loaders provide. The device side is a fixed
        ring of three staging buffers reused in rotation (a slot is rewritten only after the step that read it has
        been enqueued and an event says so): no allocator traffic inside the training loop -- fresh 50 MB tensors per
        step made the caching allocator fall back to cudaMalloc / cudaFree now and then (3 ms stalls)."""
        copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        slots = [None, None, None]  # [x, y, ready event, free event]

        def stage(batch, k):
            xin, yin = batch[0], batch[1]
            if xin.is_cuda and yin.is_cuda:  # device-resident loader: nothing to stage
                return xin.float().contiguous(), yin.float().contiguous(), None, None
            slot = slots[k]
            if slot is None or slot[0].shape != xin.shape or slot[1].shape != yin.shape:
                slot = slots[k] = [torch.empty(xin.shape, dtype=torch.float32, device=self.device),
                                   torch.empty(yin.shape, dtype=torch.float32, device=self.device),
                                   torch.cuda.Event(), None]
            with torch.cuda.stream(copy_stream):
                if slot[3] is not None:
                    copy_stream.wait_event(slot[3])  # the step that consumed this slot has been enqueued and read it
                slot[0].copy_(xin, non_blocking=True)
                slot[1].copy_(yin, non_blocking=True)
                slot[2].record(copy_stream)
            return slot[0], slot[1], slot[2], k

        it = iter(loader)
        try:
            nxt = stage(next(it), 0)
        except StopIteration:
            return
        count = 0
        while nxt is not None:
            x, y, ready, k = nxt
            count += 1
            try:
                nxt = stage(next(it), count % 3)
            except StopIteration:
                nxt = None
            if ready is not None:
                main.wait_event(ready)
            yield x, y
            if k is not None:  # the consumer has enqueued its reads of this slot on the main stream
                ev = torch.cuda.Event()
                ev.record(main)
                slots[k][3] = ev

    def _train_paired_modules(self):
        """The reference loop (model.py:598-658) over the drop-in modules: every network call is one autograd node
        executing native kernels; losses and Adam are the reference's torch objects."""
        dev = self.device
        G, D = self.generator, self.discriminator

        def lsgan(pred, target):
            return self.loss_func(pred, torch.full(pred.shape, target, dtype=torch.float32, device=dev))

        for epoch in range(self.starting_epoch, self.num_epochs + 1):
            t0 = time.time()
            losses = self.initialise_loss_storage(overall=False)
            D.train()
            G.train()
            torch.manual_seed(epoch)
            for input_stack, output_image, _ in self.train_loader:
                x = input_stack.to(dev).float()
                y = output_image.to(dev).float()
                synthetic = G(x)
                concat_real = torch.cat((x, y), 1)
                concat_synth = torch.cat((x, synthetic), 1)
                for p in D.parameters():
                    p.requires_grad = True
                self.optimizer_discriminator.zero_grad()
                d_syn = lsgan(D(concat_synth.detach()), 0.0)
                d_real = lsgan(D(concat_real), 1.0)
                ((d_syn + d_real) * 0.5).backward()
                self._allreduce_module_grads(list(D.parameters()))
                self.optimizer_discriminator.step()
                for p in D.parameters():
                    p.requires_grad = False
                self.optimizer_generator.zero_grad()
                g_adv = lsgan(D(concat_synth), 1.0)
                g_l1 = self.l1_loss(synthetic, y) * 100
                (g_adv + g_l1).backward()
                self._allreduce_module_grads(list(G.parameters()))
                self.optimizer_generator.step()
                vals = torch.stack([d_real.detach(), d_syn.detach(), g_adv.detach(), g_l1.detach()])
                if self._world() > 1:
                    dist.all_reduce(vals)
                    vals /= self._world()
                vals = vals.tolist()
                for k, v in zip(native_trainer.PairedTrainer.LOSS_KEYS, vals):
                    losses[k].append(v)
            self.scheduler_discriminator.step()
            self.scheduler_generator.step()
            self.save_results(epoch=epoch, losses=losses, epoch_start_time=t0)

    def _flush_losses(self, pending, losses, keep_last=0, keys=native_trainer.PairedTrainer.LOSS_KEYS):
        ready = pending[:len(pending) - keep_last]
        if not ready:
            return
        del pending[:len(ready)]
        vals = torch.stack(ready)
        if self._world() > 1:
            dist.all_reduce(vals)
            vals /= self._world()
        vals = vals.tolist()  # one device->host sync for the whole interval
        for row in vals:
            for k, v in zip(keys, row):
                losses[k].append(v)

    def train_cycle(self):
        """Cycle training (reference :660-758): the fused native step (fpgan.trainer.CycleTrainer -- no autograd graph,
        device-resident history buffers, gradient all-reduce per optimiser phase, CUDA-graph replay). FPG_CYCLE_MODULES=1
        selects the reference loop over the drop-in modules' autograd path instead (single process only)."""
        if os.environ.get("FPG_CYCLE_MODULES", "0") == "1":
            return self._train_cycle_modules()
        tr = self._ensure_native_cycle()
        for epoch in range(self.starting_epoch, self.num_epochs + 1):
            t0 = time.time()
            losses = self.initialise_loss_storage(overall=False)
            for net in (self.pre_to_post_generator, self.post_to_pre_generator, self.pre_discriminator,
                        self.post_discriminator):
                net.train()
            torch.manual_seed(epoch)
            lr_g = self.optimizer_generator.param_groups[0]["lr"]
            lr_d = self.optimizer_discriminator.param_groups[0]["lr"]
            pending = []
            for x, y in self._prefetch_to_device(self.train_loader):
                tr.step(x, y, lr_g=lr_g, lr_d=lr_d)
                pending.append(tr.loss_buf.clone())
                if len(pending) > self.log_interval:
                    self._flush_losses(pending, losses, keep_last=1, keys=tr.loss_keys)
            self._flush_losses(pending, losses, keys=tr.loss_keys)
            self.scheduler_generator.step()
            self.scheduler_discriminator.step()
            self.save_results(epoch=epoch, losses=losses, epoch_start_time=t0)

    def _train_cycle_modules(self):
        """The reference loop (model.py:660-758) over the drop-in modules' autograd path (kept for cross-checks)."""
        if self._world() > 1:
            raise NotImplementedError("the module-path cycle loop is single-process; unset FPG_CYCLE_MODULES for the "
                                      "data-parallel fused step")
        dev = self.device
        pre_buf, post_buf = [], []
        G_pp, G_pr = self.pre_to_post_generator, self.post_to_pre_generator
        D_pre, D_post = self.pre_discriminator, self.post_discriminator

        def lsgan(pred, target):
            return self.loss_func(pred, torch.full(pred.shape, target, dtype=torch.float32, device=dev))

        for epoch in range(self.starting_epoch, self.num_epochs + 1):
            t0 = time.time()
            losses = self.initialise_loss_storage(overall=False)
            for net in (G_pp, G_pr, D_pre, D_post):
                net.train()
            torch.manual_seed(epoch)
            for input_stack, output_image, _ in self.train_loader:
                real_pre = input_stack.to(dev).float()
                real_post = output_image.to(dev).float()
                cond = None
                if self.topography:
                    cond = real_pre[:, 3:].detach().clone()
                    real_post = torch.cat((real_post, cond), dim=1)
                synth_post, synth_pre = G_pp(real_pre), G_pr(real_post)
                if self.topography:
                    synth_post = torch.cat((synth_post, cond), dim=1)
                    synth_pre = torch.cat((synth_pre, cond), dim=1)
                rec_post, rec_pre = G_pp(synth_pre), G_pr(synth_post)
                for p in itertools.chain(D_pre.parameters(), D_post.parameters()):
                    p.requires_grad = False
                self.optimizer_generator.zero_grad()
                idt_post = idt_pre = 0
                if self.add_identity_loss:
                    idt_post = self.identity_loss(G_pp(real_post), real_post[:, :3]) * 5
                    idt_pre = self.identity_loss(G_pr(real_pre), real_pre[:, :3]) * 5
                g_post = lsgan(D_post(synth_post), 1.0)
                g_pre = lsgan(D_pre(synth_pre), 1.0)
                cyc_pre = self.cycle_loss(rec_pre, real_pre[:, :3]) * 10
                cyc_post = self.cycle_loss(rec_post, real_post[:, :3]) * 10
                (g_post + g_pre + cyc_pre + cyc_post + idt_post + idt_pre).backward()
                self.optimizer_generator.step()
                for p in itertools.chain(D_pre.parameters(), D_post.parameters()):
                    p.requires_grad = True
                self.optimizer_discriminator.zero_grad()
                sp = self.get_buffer_image(synth_pre, pre_buf)
                spo = self.get_buffer_image(synth_post, post_buf)
                d_real_pre, d_syn_pre = lsgan(D_pre(real_pre), 1.0), lsgan(D_pre(sp.detach()), 0.0)
                ((d_real_pre + d_syn_pre) * 0.5).backward()
                d_real_post, d_syn_post = lsgan(D_post(real_post), 1.0), lsgan(D_post(spo.detach()), 0.0)
                ((d_real_post + d_syn_post) * 0.5).backward()
                self.optimizer_discriminator.step()
                step_losses = [g_post, g_pre, cyc_pre, cyc_post, d_real_pre, d_real_post, d_syn_pre, d_syn_post]
                if self.add_identity_loss:
                    step_losses += [idt_post, idt_pre]
                vals = torch.stack([v.detach() for v in step_losses]).tolist()
                for k, v in zip(losses.keys(), vals):
                    losses[k].append(v)
            self.scheduler_generator.step()
            self.scheduler_discriminator.step()
            self.save_results(epoch=epoch, losses=losses, epoch_start_time=t0)
