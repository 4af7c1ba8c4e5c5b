"""Pins the oracle (oracle/gan_oracle.py) to the reference: against the committed golden vectors produced by the
unmodified reference (tests/golden/reference_vectors.json), and -- where /root/reference exists (authoring
container only) -- against the live reference modules."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gan_oracle as O  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))


def digest(params):
    return {k: {"sum": v.double().sum().item(), "abs_sum": v.double().abs().sum().item()}
            for k, v in params.items() if v.is_floating_point()}


def assert_digest(got, want, rtol, what):
    assert set(got) == set(want), f"{what}: state_dict keys differ: {set(got) ^ set(want)}"
    for k in want:
        for f in ("sum", "abs_sum"):
            a, b = got[k][f], want[k][f]
            assert abs(a - b) <= rtol * max(abs(b), 1e-3) + 1e-7, f"{what} {k}.{f}: {a} vs {b}"


def sample(t, n=64):
    f = t.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return f[idx]


@pytest.mark.parametrize("key", ["pairedattention_64", "pairedattention_256"])
def test_paired_oracle_matches_reference_golden(key):
    gold = GOLD[key]
    torch.set_num_threads(8)
    nets = O.init_model("pairedattention", "all", seed=47)
    # identical initial weights (bitwise identical draws -> digests agree to fp64 rounding)
    assert_digest(digest(nets["generator"]), gold["init"]["generator"], 1e-12, "G init")
    assert_digest(digest(nets["discriminator"]), gold["init"]["discriminator"], 1e-12, "D init")
    tr = O.PairedTrainer(nets)
    steps = len(gold["losses"]) if key.endswith("_64") else 1  # keep the 256x256 case to one CPU step
    for step in range(steps):
        x, y = O.synthetic_batch(step, gold["batch"], 9, gold["size"])
        out = tr.step(x, y)
        got = [out[k] for k in ("losses_discriminator_real", "losses_discriminator_synthetic",
                                "losses_generator_synthetic", "l1_losses_generator_synthetic")]
        for g, w in zip(got, gold["losses"][step]):
            assert abs(g - w) <= 2e-4 * abs(w) + 1e-5, f"{key} step {step}: {got} vs {gold['losses'][step]}"
    if steps == len(gold["losses"]):
        x, _ = O.synthetic_batch(0, gold["batch"], 9, gold["size"])
        with torch.no_grad():
            out, mask = O.attention_generator_forward(tr.G, x, return_mask=True)
        torch.testing.assert_close(sample(out), torch.tensor(gold["final_generator_output"]["samples"],
                                                             dtype=torch.float64), rtol=5e-3, atol=5e-4)
        torch.testing.assert_close(sample(mask), torch.tensor(gold["final_mask"]["samples"], dtype=torch.float64),
                                   rtol=5e-3, atol=5e-4)


@pytest.mark.parametrize("key,model", [("cyclegan_64", "cyclegan"), ("attentiongan_64_identity", "attentiongan")])
def test_cycle_oracle_matches_reference_golden(key, model):
    gold = GOLD[key]
    torch.set_num_threads(8)
    nets = O.init_model(model, "all", seed=47)
    for name in nets:
        assert_digest(digest(nets[name]), gold["init"][name], 1e-12, f"{name} init")
    tr = O.CycleTrainer(nets, model, add_identity_loss=gold["identity"])
    for step in range(len(gold["losses"])):
        x, y = O.synthetic_batch(step, gold["batch"], 9, gold["size"])
        out = tr.step(x, y)
        got = [out[k[4:]] for k in gold["loss_keys"]]
        for g, w in zip(got, gold["losses"][step]):
            assert abs(g - w) <= 2e-4 * abs(w) + 1e-5, f"{key} step {step}: {got} vs {gold['losses'][step]}"


def test_flood_mask_oracle_and_threshold_facts():
    fm = GOLD["flood_mask"]
    x = torch.tensor(fm["inputs"], dtype=torch.float32)
    assert O.flood_mask(x).tolist() == fm["mask"]
    # the reference expression is a step function of x switching between these two adjacent fp32 values
    assert fm["first_true_bits"] == fm["last_false_bits"] + 1 == 0x33C00001
    assert not fm["any_negative_true"]
    thr = torch.tensor([fm["last_false_bits"]], dtype=torch.int32).view(torch.float32)
    assert thr.item() == 1.5 * 2.0 ** -24
    assert O.confusion_counts(torch.tensor([1., 1, 0, 0]), torch.tensor([1., 0, 0, 1])) == [1, 1, 1, 1]


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference sources not present")
def test_oracle_forward_matches_live_reference_modules():
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_model_architectures",
                                                  "/root/reference/models/model_architectures.py")
    ref_arch = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_arch)  # the file imports only torch
    torch.manual_seed(3)
    x = torch.rand(1, 9, 64, 64) * 2 - 1
    for cls, plan, fwd in ((ref_arch.PairedAttentionGenerator, O.attention_generator_plan,
                            O.attention_generator_forward),
                           (ref_arch.CycleGANGenerator, O.cyclegan_generator_plan, O.cyclegan_generator_forward)):
        m = cls(9)
        p = {k: v.detach().clone() for k, v in m.state_dict().items()}
        assert list(p) == [n + s for n, shape, b, kind in plan(9) for s in ([".weight", ".bias"] if b else [".weight"])]
        with torch.no_grad():
            torch.testing.assert_close(fwd(p, x), m(x), rtol=1e-5, atol=1e-6)
    d = ref_arch.PairedAttentionDiscriminator(9)
    p = {k: v.detach().clone() for k, v in d.state_dict().items()}
    xd = torch.rand(1, 12, 64, 64)
    with torch.no_grad():
        torch.testing.assert_close(O.patchgan_forward(p, xd), d(xd), rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference sources not present")
def test_pix2pix_oracle_matches_live_reference_modules_at_batch_2():
    """BatchNorm over a batch of TWO (the goldens step at batch 1) and dropout from the same seeded RNG: the oracle's
    Pix2Pix U-Net and BatchNorm PatchGAN against the live reference modules in training mode, running statistics
    included."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_model_architectures",
                                                  "/root/reference/models/model_architectures.py")
    ref_arch = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_arch)
    torch.manual_seed(5)
    g, d = ref_arch.Pix2PixGenerator(9), ref_arch.Pix2PixDiscriminator(9)
    g.train()
    d.train()
    pg = {k: v.detach().clone() for k, v in g.state_dict().items()}
    pd = {k: v.detach().clone() for k, v in d.state_dict().items()}
    x = torch.rand(2, 9, 256, 256) * 2 - 1
    xd = torch.rand(2, 12, 256, 256) * 2 - 1
    with torch.no_grad():
        torch.manual_seed(11)
        want = g(x)
        torch.manual_seed(11)
        got = O.pix2pix_generator_forward(pg, x)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(O.patchgan_bn_forward(pd, xd), d(xd), rtol=1e-4, atol=1e-5)
    for params, module in ((pg, g), (pd, d)):
        sd = module.state_dict()
        for k, v in params.items():
            if "running_" in k:
                torch.testing.assert_close(v, sd[k], rtol=1e-4, atol=1e-6, msg=k)
            if "num_batches" in k:
                assert int(v) == int(sd[k]) == 1, k


def test_pix2pix_oracle_matches_reference_golden():
    """Pix2Pix (BASELINE.json configs[0]): 8-level U-Net with BatchNorm in training mode, in-place activations feeding
    the skip connections and dropout from the torch RNG seeded per epoch (model.py:609)."""
    gold = GOLD["pix2pix_256"]
    torch.set_num_threads(8)
    nets = O.init_model("pix2pix", "all", seed=47)
    assert len(nets["generator"]) == 82
    assert_digest(digest(nets["generator"]), gold["init"]["generator"], 1e-12, "G init")
    assert_digest(digest(nets["discriminator"]), gold["init"]["discriminator"], 1e-12, "D init")
    tr = O.PairedTrainer(nets, "pix2pix")
    for step in range(len(gold["losses"])):
        x, y = O.synthetic_batch(step, gold["batch"], 9, gold["size"])
        torch.manual_seed(gold["epoch_seed"])
        out = tr.step(x, y)
        got = [out[k] for k in ("losses_discriminator_real", "losses_discriminator_synthetic",
                                "losses_generator_synthetic", "l1_losses_generator_synthetic")]
        for g, w in zip(got, gold["losses"][step]):
            assert abs(g - w) <= 2e-4 * abs(w) + 1e-5, f"pix2pix step {step}: {got} vs {gold['losses'][step]}"
    # parameters AND BatchNorm running statistics after the two steps
    assert_digest(digest(nets["generator"]), gold["final"]["generator"], 2e-3, "G final")
    assert_digest(digest(nets["discriminator"]), gold["final"]["discriminator"], 2e-3, "D final")


def test_unet_oracle_matches_reference_golden():
    """Segmentation U-Net as used by calculate_metrics (model.py:380-418, BASELINE.json configs[4]): same initial
    weights, logits, BatchNorm buffer side effects, and -- integer work -- identical flood masks and confusion counts."""
    gold = GOLD["unet_64"]
    torch.set_num_threads(8)
    p = O.init_unet(47)
    assert_digest(digest(p), gold["init"], 1e-12, "U-Net init")
    g = torch.Generator().manual_seed(2000)
    gen = torch.rand(gold["batch"], 3, gold["size"], gold["size"], generator=g) * 2 - 1
    truth = torch.rand(gold["batch"], 3, gold["size"], gold["size"], generator=g) * 2 - 1
    with torch.no_grad():
        lg = O.unet_forward(p, torch.clamp((gen + 1) * 0.5, min=0, max=1))
        lt = O.unet_forward(p, torch.clamp((truth + 1) * 0.5, min=0, max=1))
    torch.testing.assert_close(sample(lg), torch.tensor(gold["logits_generated"]["samples"], dtype=torch.float64),
                               rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(sample(lt), torch.tensor(gold["logits_truth"]["samples"], dtype=torch.float64),
                               rtol=1e-4, atol=1e-6)
    mo, mt = O.flood_mask(lg), O.flood_mask(lt)
    assert float(mo.sum()) == gold["mask_generated_sum"] and float(mt.sum()) == gold["mask_truth_sum"]
    assert O.confusion_counts(mo.flatten(), mt.flatten()) == gold["confusion_tp_fp_tn_fn"]
    assert_digest(digest(p), gold["final"], 1e-5, "U-Net buffers after two forward passes")
