"""Generates tests/golden/reference_vectors.json by running the UNMODIFIED reference (imported from /root/reference,
which exists only in the authoring container) on seeded synthetic inputs. The committed JSON is what pins the oracle
(oracle/gan_oracle.py) on machines where the reference is absent.

    python tests/golden/make_golden.py        # rewrites reference_vectors.json

Recipe: tifffile / matplotlib / torchmetrics are not installed, so empty stub modules are
injected for those imports; Model.train_loader is replaced by a list of synthetic batches; nothing in the
reference's maths is touched.
"""
import json
import os
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    def stub(name, attrs=()):
        m = types.ModuleType(name)
        sys.modules[name] = m
        for a in attrs:
            setattr(m, a, type(a, (), {}))
        return m

    stub("tifffile")
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("torchmetrics")
    stub("torchmetrics.regression", ["MeanSquaredError"])
    stub("torchmetrics.image", ["PeakSignalNoiseRatio", "MultiScaleStructuralSimilarityIndexMeasure",
                                "StructuralSimilarityIndexMeasure"])
    stub("torchmetrics.image.lpip", ["LearnedPerceptualImagePatchSimilarity"])
    stub("torchmetrics.classification", ["BinaryAccuracy", "BinaryF1Score", "BinaryPrecision", "BinaryRecall"])
    sys.path.insert(0, REF)
    os.chdir(REF)  # data.py reads metadata/dataset_split.csv relative to the cwd
    from models import model as ref_model
    return ref_model


def batches(steps, batch, size, channels=9):
    out = []
    for step in range(steps):
        g = torch.Generator().manual_seed(1000 + step)
        x = torch.rand(batch, channels, size, size, generator=g) * 2 - 1
        y = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
        out.append((x, y, ("synthetic",)))
    return out


def sample(t, n=64):
    """n deterministic samples of a tensor (flattened, evenly spaced) + its sum and abs-sum"""
    f = t.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return {"samples": f[idx].tolist(), "sum": f.sum().item(), "abs_sum": f.abs().sum().item(), "numel": f.numel()}


def param_digest(module):
    return {k: {"sum": v.double().sum().item(), "abs_sum": v.double().abs().sum().item()}
            for k, v in module.state_dict().items() if v.is_floating_point()}


def run_paired(ref_model, size, batch, steps, model="pairedattention"):
    M = ref_model.Model(model=model, dataset_subset="usa", dataset_dem="same", data_path="/tmp/none",
                        num_epochs=200, topography="all", resize=512, crop=4, training_model=True, seed=47)
    init = {"generator": param_digest(M.generator), "discriminator": param_digest(M.discriminator)}
    per_step = []
    data = batches(steps, batch, size)
    for b in data:  # one "epoch" per step so that per-step losses are observable through the public API.
        # num_epochs stays 200 so that lambda_rule (model.py:175-181) keeps the lr at 2e-4 for all of these steps
        M.train_loader = [b]
        M.starting_epoch = M.num_epochs
        M.all_losses = M.initialise_loss_storage(overall=True)
        # train_paired reseeds torch with the epoch number (model.py:609); harmless: no RNG use in this model
        M.train_paired()
        per_step.append([float(M.all_losses[k][-1]) for k in
                         ("all_losses_discriminator_real", "all_losses_discriminator_synthetic",
                          "all_losses_generator_synthetic", "all_l1_losses_generator_synthetic")])
    res = {"size": size, "batch": batch, "losses": per_step, "init": init, "epoch_seed": M.num_epochs,
           "final": {"generator": param_digest(M.generator), "discriminator": param_digest(M.discriminator)}}
    if model == "pairedattention":
        with torch.no_grad():
            out = M.generator(data[0][0])
        res["final_generator_output"] = sample(out)
        res["final_mask"] = sample(M.generator.last_attention_mask)
    return res


def run_cycle(ref_model, name, size, batch, steps, identity):
    M = ref_model.Model(model=name, dataset_subset="usa", dataset_dem="same", data_path="/tmp/none", num_epochs=200,
                        topography="all", resize=512, crop=4, training_model=True, seed=47,
                        add_identity_loss=identity)
    init = {k: param_digest(getattr(M, k)) for k in
            ("pre_to_post_generator", "post_to_pre_generator", "pre_discriminator", "post_discriminator")}
    per_step = []
    data = batches(steps, batch, size)
    keys = None
    for b in data:
        M.train_loader = [b]
        M.starting_epoch = M.num_epochs
        M.all_losses = M.initialise_loss_storage(overall=True)
        M.train_cycle()
        keys = list(M.all_losses.keys())
        per_step.append([float(M.all_losses[k][-1]) for k in keys])
    with torch.no_grad():
        out = M.pre_to_post_generator(data[0][0])
    return {"size": size, "batch": batch, "identity": identity, "loss_keys": keys, "losses": per_step, "init": init,
            "final_generator_output": sample(out)}


def run_unet(size, batch):
    """The segmentation U-Net exactly as calculate_metrics uses it (model.py:380-400): constructed and initialised like
    SegmentationModel.__init__ (segmentation_model.py:55), never switched to eval mode, applied to the generated and the
    ground-truth image rescaled to [0, 1]; masks by (sigmoid > 0.5)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_arch", os.path.join(REF, "models", "model_architectures.py"))
    arch = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(arch)

    def init(m):  # segmentation_model.py:73-84
        name = m.__class__.__name__
        if hasattr(m, "weight") and (name.find("Conv") != -1 or name.find("Linear") != -1):
            torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
            if hasattr(m, "bias") and m.bias is not None:
                torch.nn.init.constant_(m.bias.data, 0.0)
        elif name.find("BatchNorm2d") != -1:
            torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
            torch.nn.init.constant_(m.bias.data, 0.0)

    torch.manual_seed(47)
    net = arch.UNet().apply(init)
    init_digest = param_digest(net)
    g = torch.Generator().manual_seed(2000)
    generated = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    truth = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    with torch.no_grad():
        gt = torch.clamp((truth + 1) * 0.5, min=0, max=1)
        gen = torch.clamp((generated + 1) * 0.5, min=0, max=1)
        logits_gen = net(gen)
        out_mask = (torch.sigmoid(logits_gen) > 0.5).float()
        logits_gt = net(gt)
        true_mask = (torch.sigmoid(logits_gt) > 0.5).float()
    p, t = out_mask.flatten() > 0.5, true_mask.flatten() > 0.5
    counts = [int((p & t).sum()), int((p & ~t).sum()), int((~p & ~t).sum()), int((~p & t).sum())]
    return {"size": size, "batch": batch, "init": init_digest, "logits_generated": sample(logits_gen),
            "logits_truth": sample(logits_gt), "mask_generated_sum": float(out_mask.sum()),
            "mask_truth_sum": float(true_mask.sum()), "confusion_tp_fp_tn_fn": counts,
            "final": param_digest(net)}


def flood_mask_facts():
    """Exhaustive scan of (sigmoid(x) > 0.5) over every positive fp32 value: the expression is a step function."""
    first_true = last_false = None
    chunk = 1 << 26
    for start in range(0, 0x7F800001, chunk):
        bits = torch.arange(start, min(start + chunk, 0x7F800001), dtype=torch.int32)
        m = torch.sigmoid(bits.view(torch.float32)) > 0.5
        if m.any() and first_true is None:
            first_true = int(bits[m][0])
        if (~m).any():
            last_false = int(bits[~m][-1])
    neg = torch.arange(-(1 << 31), -(1 << 31) + 0x7F800001, 1 << 3, dtype=torch.int64).to(torch.int32)
    any_neg = bool((torch.sigmoid(neg.view(torch.float32)) > 0.5).any())
    g = torch.Generator().manual_seed(5)
    x = torch.cat([torch.randn(48, generator=g) * 2,
                   torch.tensor([0.0, -0.0, 5.9e-8, 8.9e-8, 8.940696716308594e-08, 8.94069742685133e-08, 1.2e-7])])
    return {"first_true_bits": first_true, "last_false_bits": last_false, "any_negative_true": any_neg,
            "inputs": x.tolist(), "mask": (torch.sigmoid(x) > 0.5).float().tolist()}


def main():
    torch.set_num_threads(8)
    ref_model = import_reference()
    out = {"torch": torch.__version__, "note": "outputs of the unmodified reference, CPU fp32"}
    out["pairedattention_64"] = run_paired(ref_model, 64, 2, 3)
    out["pairedattention_256"] = run_paired(ref_model, 256, 1, 2)
    # Pix2Pix: BatchNorm in training mode + dropout drawn from the global RNG that train_paired seeds with the epoch
    # number (model.py:609; every step here is its own "epoch" 200). 256x256 is the smallest input of the 8-level U-Net
    out["pix2pix_256"] = run_paired(ref_model, 256, 1, 2, model="pix2pix")
    out["cyclegan_64"] = run_cycle(ref_model, "cyclegan", 64, 1, 2, False)
    out["attentiongan_64_identity"] = run_cycle(ref_model, "attentiongan", 64, 1, 2, True)
    out["unet_64"] = run_unet(64, 4)
    out["flood_mask"] = flood_mask_facts()
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote reference_vectors.json")


if __name__ == "__main__":
    main()
