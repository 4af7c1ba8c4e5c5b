"""GPU parity of the drop-in modules and the fused training step against the oracle (oracle/gan_oracle.py, a CPU
fp32 restatement pinned to the reference by tests/test_oracle_cpu.py), on identical seeded inputs and weights.

Tolerances (bf16 bound of north_star; measured values are printed with `-s`):
  * per-step losses: rtol 2e-2 (LOSS_RTOL); measured 1e-3 .. 2e-3.
  * forward tensors (generator output, attention mask, PatchGAN logits): RMS error relative to the RMS of the oracle
    tensor < OUT_TOL = 2e-2 (north_star's bf16 bound). An element-wise rtol is meaningless for values near zero.
    Round 1 stored every tensor as bf16 and measured 2.25e-2 = what a CPU simulation of the four rounding sites
    (weights, conv inputs, conv outputs, residual stream) predicts (tests/sim_bf16_rounding.py: 2.27e-2). The two
    sites no tensor-core instruction reads -- pre-norm conv outputs and the residual skip stream -- are now stored as
    fp16 (same bytes, 3 more mantissa bits): simulated 1.75e-2, bf16 operand rounding alone being 1.74e-2.
  * parameter / input gradients: RMS-relative error < GRAD_TOL = 0.3. A forward deviation eps flips the ReLU /
    LeakyReLU mask of a fraction ~0.8*eps of the elements (those whose pre-activation lies within the noise of zero);
    each flip is a 100% element error, so one activation stage alone contributes sqrt(0.8*eps) ~ 13% for eps = 2%.
    This is unbiased noise, not drift: the losses agree to 2e-3.
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

LOSS_RTOL = 2e-2
OUT_TOL = 2e-2
GRAD_TOL = 0.3


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def rel_rms(got, ref):
    ref = ref.float().cpu()
    got = got.float().cpu()
    return ((got - ref).pow(2).mean().sqrt() / (ref.pow(2).mean().sqrt() + 1e-12)).item()


def make_pair(seed=47):
    """oracle parameter dicts + drop-in modules holding the same weights"""
    from oracle import gan_oracle as O
    from models import model_architectures as A
    nets = O.init_model("pairedattention", "all", seed=seed)
    torch.manual_seed(seed)
    G = A.PairedAttentionGenerator(9)
    D = A.PairedAttentionDiscriminator(9)
    G.load_state_dict(nets["generator"])
    D.load_state_dict(nets["discriminator"])
    return O, nets, G.cuda(), D.cuda()


def test_drop_in_constructors_initialise_like_the_reference():
    """Model(seed=47) wiring: same construction order and RNG consumption as the reference (checked against the
    golden parameter digests of the unmodified reference)."""
    import json
    from models import model as M
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))["pairedattention_64"]
    m = M.Model(model="PairedAttention", topography="all", num_epochs=200, seed=47, training_model=True)
    for net, key in ((m.generator, "generator"), (m.discriminator, "discriminator")):
        sd = net.state_dict()
        assert set(sd) == set(gold["init"][key])
        for k, v in sd.items():
            want = gold["init"][key][k]
            assert abs(v.double().sum().item() - want["sum"]) <= 1e-6 * max(1.0, abs(want["abs_sum"]))
            assert abs(v.double().abs().sum().item() - want["abs_sum"]) <= 1e-6 * max(1.0, want["abs_sum"])


@pytest.mark.parametrize("size,batch", [(64, 2), (256, 1)])
def test_generator_forward_backward_matches_oracle(size, batch):
    O, nets, G, D = make_pair()
    x, _ = O.synthetic_batch(0, batch, 9, size)
    wgt = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(1))
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in nets["generator"].items()}
    xr = x.clone().requires_grad_(True)
    out_ref, mask_ref = O.attention_generator_forward(p, xr, return_mask=True)
    names = [k for k in p]
    grads_ref = torch.autograd.grad((out_ref * wgt).sum(), [p[k] for k in names] + [xr])
    xg = x.cuda().requires_grad_(True)
    out = G(xg)
    e_out = rel_rms(out, out_ref.detach())
    assert e_out < OUT_TOL, e_out
    assert rel_rms(G.last_attention_mask, mask_ref.detach()) < OUT_TOL
    (out * wgt.cuda()).sum().backward()
    gp = dict(G.named_parameters())
    worst = 0.0
    for k, gref in zip(names, grads_ref[:-1]):
        if k.endswith(".bias") and not k.startswith("deconv3"):
            # bias before an InstanceNorm: mathematically zero gradient (reference gets ~1e-9 rounding noise)
            assert gp[k].grad.abs().max().item() == 0.0
            continue
        e = rel_rms(gp[k].grad, gref)
        worst = max(worst, e)
        if os.environ.get("FPG_VERBOSE_PARITY"):
            print(f"[parity] generator grad {k}: rel-rms err {e:.4f}")
    for k, gref in zip(names, grads_ref[:-1]):
        if k.endswith(".bias") and not k.startswith("deconv3"):
            continue
        e = rel_rms(gp[k].grad, gref)
        assert e < GRAD_TOL, f"grad {k}: rel rms err {e}"
    e_dx = rel_rms(xg.grad, grads_ref[-1])
    assert e_dx < GRAD_TOL, e_dx
    print(f"\n[parity] generator {size}x{size} B={batch}: output rel-rms err {e_out:.4f}, worst parameter-gradient "
          f"rel-rms err {worst:.4f}, input-gradient err {e_dx:.4f}")


def test_discriminator_forward_backward_matches_oracle():
    O, nets, G, D = make_pair()
    x = torch.rand(2, 12, 256, 256, generator=torch.Generator().manual_seed(2)) * 2 - 1
    p = {k: v.clone().requires_grad_(True) for k, v in nets["discriminator"].items()}
    xr = x.clone().requires_grad_(True)
    out_ref = O.patchgan_forward(p, xr)
    loss_ref = torch.nn.functional.mse_loss(out_ref, torch.ones_like(out_ref))
    names = list(p)
    grads_ref = torch.autograd.grad(loss_ref, [p[k] for k in names] + [xr])
    xg = x.cuda().requires_grad_(True)
    out = D(xg)
    assert out.shape == (2, 1, 30, 30)
    e_out = rel_rms(out, out_ref.detach())
    assert e_out < OUT_TOL, e_out
    torch.nn.functional.mse_loss(out, torch.ones_like(out)).backward()
    gp = dict(D.named_parameters())
    for k, gref in zip(names, grads_ref[:-1]):
        if k in ("model.2.bias", "model.5.bias", "model.8.bias"):
            assert gp[k].grad.abs().max().item() == 0.0
            continue
        e = rel_rms(gp[k].grad, gref)
        print(f"[parity] discriminator grad {k}: rel-rms err {e:.4f}")
        assert e < GRAD_TOL, f"grad {k}: {e}"
    assert rel_rms(xg.grad, grads_ref[-1]) < GRAD_TOL
    # frozen discriminator (generator phase, model.py:636-637): only the input gradient is produced
    for q in D.parameters():
        q.requires_grad = False
        q.grad = None
    xg2 = x.cuda().requires_grad_(True)
    torch.nn.functional.mse_loss(D(xg2), torch.ones_like(out)).backward()
    assert rel_rms(xg2.grad, grads_ref[-1]) < GRAD_TOL
    assert all(q.grad is None for q in D.parameters())


@pytest.mark.parametrize("size,batch,steps", [(64, 2, 3), (256, 1, 2)])
def test_fused_paired_step_matches_oracle_and_reference_golden(size, batch, steps):
    """train_paired (model.py:611-651): per-step losses vs the oracle AND vs the golden losses of the unmodified
    reference, free-running for a few steps."""
    import json
    from fpgan.trainer import PairedTrainer
    O, nets, G, D = make_pair()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))[f"pairedattention_{size}"]
    assert gold["batch"] == batch
    otr = O.PairedTrainer(nets)
    tr = PairedTrainer(G, D)
    keys = PairedTrainer.LOSS_KEYS
    for step in range(steps):
        x, y = O.synthetic_batch(step, batch, 9, size)
        ref = otr.step(x, y)
        synth = tr.step(x.cuda(), y.cuda())
        got = tr.losses()
        e_out = rel_rms(synth, ref["synthetic"])
        print(f"\n[parity] step {step}: output rel-rms err {e_out:.4f}; losses " +
              ", ".join(f"{got[k]:.5f}/{ref[k]:.5f}" for k in keys))
        if step == 0:  # later steps are free-running: Adam's first updates are ~lr*sign(g), so sign flips of small
            assert e_out < OUT_TOL, e_out  # gradients separate the weight trajectories; the losses still agree
        for i, k in enumerate(keys):
            assert abs(got[k] - ref[k]) <= LOSS_RTOL * abs(ref[k]) + 1e-4, f"step {step} {k}: {got[k]} vs oracle {ref[k]}"
            w = gold["losses"][step][i]
            assert abs(got[k] - w) <= LOSS_RTOL * abs(w) + 1e-4, f"step {step} {k}: {got[k]} vs reference {w}"
    # parameters after the updates: state_dict round-trips through the flat buffers
    sd = G.state_dict()
    e = rel_rms(sd["resnet_blocks.4.conv1.weight"], otr.G["resnet_blocks.4.conv1.weight"])
    assert e < 0.1, f"weights drifted: {e}"


def test_fused_paired_step_at_the_benchmarked_configuration():
    """ONE fused train_paired step at BASELINE configs[1] -- batch 16, 256x256, 9 channels -- against the oracle's
    fp32 CPU step on the same inputs: losses rtol 2e-2, generated images < OUT_TOL. At this batch the planner picks
    kernels the small cases never launch (2-CTA fprop, row-stationary 7x7 heads at full width, CTA-pair wgrad); the
    kernel names recorded by CUPTI during the step are checked so that the parity claim covers them."""
    from fpgan.trainer import PairedTrainer
    O, nets, G, D = make_pair()
    batch, size = 16, 256
    x, y = O.synthetic_batch(0, batch, 9, size)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = O.PairedTrainer(nets).step(x, y)
    tr = PairedTrainer(G, D)
    names = set()
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            synth = tr.step(x.cuda(), y.cuda())
            torch.cuda.synchronize()
        names = {e.name for e in prof.events() if "fpg::" in e.name}
    except Exception as exc:  # CUPTI not usable on this box: the plan flags below still pin the kernel choice
        print(f"[parity] kernel-name capture unavailable ({exc})")
        synth = tr.step(x.cuda(), y.cuda())
    got = tr.losses()
    e_out = rel_rms(synth, ref["synthetic"])
    print(f"\n[parity] B=16 256x256 step: output rel-rms err {e_out:.4f}; losses " +
          ", ".join(f"{got[k]:.5f}/{ref[k]:.5f}" for k in PairedTrainer.LOSS_KEYS))
    assert e_out < OUT_TOL, e_out
    for k in PairedTrainer.LOSS_KEYS:
        assert abs(got[k] - ref[k]) <= LOSS_RTOL * abs(ref[k]) + 1e-4, f"{k}: {got[k]} vs oracle {ref[k]}"
    if names:
        for want in ("igemm_fprop2_kernel", "igemm_rows_kernel", "igemm_wgrad2_kernel", "igemm_fprop_kernel",
                     "in_apply_ring_kernel", "blend_fwd_kernel", "adam"):
            assert any(want in n for n in names), f"{want} was not launched at the benchmarked configuration: {sorted(names)}"
    # the same facts from the planners (pure host code): residual conv -> CTA pair, 7x7 head -> row-stationary
    import ctypes as C
    from fpgan import lib as L, ops
    spec = tr.G.layers["resnet_blocks.0.conv1"].spec
    xb = ops.ActBuf(batch, 64, 64, 256, halo=1, zero=False)
    yb = ops.ActBuf(batch, 64, 64, 256, zero=False)
    d = L.FpropDesc()
    L.call("fpg_conv2d_fprop_plan", xb.ref(), ops._ptr(spec.w_fprop), None, 0, spec.gref(), yb.ref(),
           L.load().fpg_sm_count(), C.byref(d))
    assert d.cta_pair == 1
    head = tr.G.layers["deconv3_content"].spec
    vb = ops.ActBuf(batch, 256, 256, 64, halo=3, zero=False)
    cb = ops.ActBuf(batch, 256, 256, 32, fp32=True, zero=False)
    rd = L.RowsDesc()
    assert L.load().fpg_conv2d_rows_plan(vb.ref(), ops._ptr(head.w_fprop), None, 0, head.gref(), cb.ref(), 0,
                                         L.load().fpg_sm_count(), C.byref(rd)) == 0


def test_model_train_paired_api_and_checkpoint_roundtrip(tmp_path):
    """Model(**kwargs).train_paired() with an injected loader, checkpoint written in the reference's dict layout and
    resumed through load_pretrained_model."""
    from models import model as M
    from models.data import SyntheticLoader
    m = M.Model(model="PairedAttention", topography="all", num_epochs=1, seed=47, data_path=str(tmp_path),
                save_model_interval=1, log_interval=2)
    m.train_loader = SyntheticLoader(steps=3, batch=2, size=64)
    m.train_paired()
    assert len(m.all_losses["all_losses_discriminator_real"]) == 1
    files = list((tmp_path / "models").glob("PairedAttention_*.pth.tar"))
    assert len(files) == 1
    ck = torch.load(files[0], weights_only=False)
    for key in ("model", "starting_epoch", "num_epochs", "topography", "optimizer_generator",
                "optimizer_discriminator", "scheduler_generator", "scheduler_discriminator", "all_losses",
                "add_identity_loss", "generator", "discriminator"):
        assert key in ck
    assert ck["starting_epoch"] == 2 and ck["model"] == "pairedattention"
    assert len(ck["optimizer_generator"]["state"]) == 54 and len(ck["generator"]) == 54
    m2 = M.Model(load_pretrained_model=True, pretrained_model_path=str(files[0]), data_path=str(tmp_path))
    assert m2.starting_epoch == 2
    for k, v in m.generator.state_dict().items():
        assert torch.equal(v, m2.generator.state_dict()[k])


@pytest.mark.parametrize("size,batch", [(64, 2), (256, 1)])
def test_fp32_parity_mode_matches_reference_to_1e4(size, batch):
    """north_star: "fp32 rtol 1e-4". The 3 x bf16-split mode (fpgan/fp32mode.py: every conv = hi*hi + lo*hi + hi*lo on
    the tcgen05 kernels, everything else fp32) against the oracle's fp32 CPU graph on identical inputs and weights:
    generator output, attention mask, PatchGAN logits and the four losses of the step."""
    from fpgan import fp32mode
    O, nets, G, D = make_pair()
    x, y = O.synthetic_batch(0, batch, 9, size)
    with torch.no_grad():
        out_ref, mask_ref = O.attention_generator_forward(nets["generator"], x, return_mask=True)
        lg_syn = O.patchgan_forward(nets["discriminator"], torch.cat((x, out_ref), 1))
        lg_real = O.patchgan_forward(nets["discriminator"], torch.cat((x, y), 1))
        ref = {"losses_discriminator_real": F.mse_loss(lg_real, torch.ones_like(lg_real)).item(),
               "losses_discriminator_synthetic": F.mse_loss(lg_syn, torch.zeros_like(lg_syn)).item(),
               "losses_generator_synthetic": F.mse_loss(lg_syn, torch.ones_like(lg_syn)).item(),
               "l1_losses_generator_synthetic": 100 * F.l1_loss(out_ref, y).item()}
    sg = fp32mode.SplitAttentionGenerator(G)
    out, mask = sg.forward(x.cuda())
    e_out, e_mask = rel_rms(out, out_ref), rel_rms(mask, mask_ref)
    worst = ((out.cpu() - out_ref).abs() / (out_ref.abs() + out_ref.abs().mean())).max().item()
    synth, got, logits = fp32mode.paired_forward_losses(G, D, x.cuda(), y.cuda())
    e_log = rel_rms(logits.to_nchw(1), lg_syn)
    print(f"\n[parity fp32 mode] {size}x{size} B={batch}: output rel-rms {e_out:.2e} (worst element {worst:.2e}), "
          f"mask {e_mask:.2e}, logits {e_log:.2e}; losses " + ", ".join(f"{got[k]:.6f}/{ref[k]:.6f}" for k in ref))
    assert e_out < 1e-4 and e_mask < 1e-4 and e_log < 1e-4, (e_out, e_mask, e_log)
    assert worst < 1e-3
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-4 * abs(ref[k]), f"{k}: {got[k]} vs {ref[k]}"


def test_cyclegan_generator_forward_backward_matches_oracle():
    """CycleGANGenerator (model_architectures.py:91-134) at the tensor level: output, parameter gradients and input
    gradient of the drop-in module against the oracle's fp32 graph."""
    from oracle import gan_oracle as O
    from models import model_architectures as A
    nets = O.init_model("cyclegan", "all", seed=47)
    G = A.CycleGANGenerator(9)
    G.load_state_dict(nets["pre_to_post_generator"])
    G = G.cuda()
    x, _ = O.synthetic_batch(0, 2, 9, 64)
    wgt = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in nets["pre_to_post_generator"].items()}
    xr = x.clone().requires_grad_(True)
    out_ref = O.cyclegan_generator_forward(p, xr)
    names = list(p)
    grads_ref = torch.autograd.grad((out_ref * wgt).sum(), [p[k] for k in names] + [xr])
    xg = x.cuda().requires_grad_(True)
    out = G(xg)
    e_out = rel_rms(out, out_ref.detach())
    assert e_out < OUT_TOL, e_out
    (out * wgt.cuda()).sum().backward()
    gp = dict(G.named_parameters())
    worst = 0.0
    for k, gref in zip(names, grads_ref[:-1]):
        if k.endswith(".bias") and k != "model.26.bias":  # bias before an InstanceNorm: exactly zero gradient here
            assert gp[k].grad.abs().max().item() == 0.0
            continue
        e = rel_rms(gp[k].grad, gref)
        worst = max(worst, e)
        assert e < GRAD_TOL, f"grad {k}: rel rms err {e}"
    e_dx = rel_rms(xg.grad, grads_ref[-1])
    assert e_dx < GRAD_TOL, e_dx
    print(f"\n[parity] CycleGAN generator 64x64 B=2: output rel-rms err {e_out:.4f}, worst parameter-gradient "
          f"rel-rms err {worst:.4f}, input-gradient err {e_dx:.4f}")


def _cycle_pair(model, identity, seed=47):
    from oracle import gan_oracle as O
    from models import model_architectures as A
    nets = O.init_model(model, "all", seed=seed)
    gen = {"cyclegan": A.CycleGANGenerator, "attentiongan": A.AttentionGANGenerator}[model]
    dis = {"cyclegan": A.CycleGANDiscriminator, "attentiongan": A.AttentionGANDiscriminator}[model]
    mods = {}
    for key, cls in (("pre_to_post_generator", gen), ("post_to_pre_generator", gen), ("pre_discriminator", dis),
                     ("post_discriminator", dis)):
        m = cls(9)
        m.load_state_dict(nets[key])
        mods[key] = m.cuda()
    return O, nets, mods


@pytest.mark.parametrize("model,identity", [("cyclegan", False), ("attentiongan", True)])
def test_fused_cycle_step_matches_oracle_teacher_forced(model, identity):
    """CycleTrainer (fused train_cycle, model.py:678-739) against the oracle's fp32 step, TEACHER-FORCED: before every
    step the native trainer adopts the oracle's weights and Adam moments, so each step -- eager steps 0-1, the captured
    step 2 and graph replays afterwards -- is an independent check of one full iteration: every loss within rtol 2e-2,
    both synthetic images within OUT_TOL."""
    from fpgan.trainer import CycleTrainer
    O, nets, mods = _cycle_pair(model, identity)
    otr = O.CycleTrainer(nets, model, add_identity_loss=identity, py_seed=3)
    tr = CycleTrainer(mods["pre_to_post_generator"], mods["post_to_pre_generator"], mods["pre_discriminator"],
                      mods["post_discriminator"], add_identity_loss=identity)

    def adopt(fp, dicts, adam):
        plist = [v for d in dicts for v in d.values() if v.is_floating_point() and v.dim() > 0]
        assert len(plist) == len(fp.named)
        for (name, p), src, m, v in zip(fp.named, plist, adam.m, adam.v):
            off, k = fp.offsets[name]
            assert p.shape == src.shape, name
            fp.flat[off:off + k].copy_(src.reshape(-1))
            fp.m[off:off + k].copy_(m.reshape(-1))
            fp.v[off:off + k].copy_(v.reshape(-1))
        fp.steps = adam.t

    worst = 0.0
    for step in range(5):
        adopt(tr.gp, (otr.G_pp, otr.G_pr), otr.opt_g)
        adopt(tr.dp, (otr.D_post, otr.D_pre), otr.opt_d)
        for net in (tr.Gpp, tr.Gpr, tr.Dpost, tr.Dpre):
            net.repack(force=True)
        x, y = O.synthetic_batch(step, 1, 9, 64)
        with torch.no_grad():
            sp_ref = O.GENERATOR_FORWARD[model](otr.G_pp, x)
            sr_ref = O.GENERATOR_FORWARD[model](otr.G_pr, torch.cat((y, x[:, 3:]), 1))
        ref = otr.step(x, y)
        sp, sr = tr.step(x.cuda(), y.cuda())
        got = tr.losses()
        e1, e2 = rel_rms(sp, sp_ref), rel_rms(sr, sr_ref)
        print(f"\n[parity] {model} cycle step {step}: synthetic post/pre rel-rms err {e1:.4f}/{e2:.4f}; losses " +
              ", ".join(f"{got[k]:.5f}/{ref[k]:.5f}" for k in tr.loss_keys))
        assert e1 < OUT_TOL and e2 < OUT_TOL, (e1, e2)
        for k in tr.loss_keys:
            err = abs(got[k] - ref[k]) / (abs(ref[k]) + 1e-6)
            worst = max(worst, err)
            assert abs(got[k] - ref[k]) <= LOSS_RTOL * abs(ref[k]) + 1e-4, f"step {step} {k}: {got[k]} vs {ref[k]}"
    assert tr._graphs, "the step was never captured"
    print(f"[parity] {model} cycle: worst loss rel err over 5 teacher-forced steps {worst:.5f}")


def test_history_exchange_kernel_follows_the_reference_buffer():
    """fpg_history_exchange + _History == get_buffer_image (model.py:275-294) over 140 steps, past the fill phase"""
    import random
    from fpgan import ops
    from fpgan.trainer import _History
    from models import model as M
    shape = (2, 4, 4, 16)
    pool = torch.zeros((50,) + shape, dtype=torch.bfloat16, device="cuda")
    ctrl = torch.zeros(2, dtype=torch.int32, device="cuda")
    out = torch.empty(shape, dtype=torch.bfloat16, device="cuda")
    random.seed(11)
    hist, got = _History(), []
    for i in range(140):
        ctrl.copy_(torch.tensor(hist.decide(), dtype=torch.int32))
        cur = torch.full(shape, float(i), dtype=torch.bfloat16, device="cuda")
        ops.history_exchange(cur, pool, ctrl, out)
        assert (out == out.flatten()[0]).all()
        got.append(int(out.flatten()[0].item()))
    random.seed(11)
    buf, want = [], []
    for i in range(140):
        want.append(int(M.Model.get_buffer_image(None, torch.full((1,), float(i)), buf).item()))
    assert got == want and [int(p.flatten()[0].item()) for p in pool] == [int(b.item()) for b in buf]
    assert any(w != i for i, w in enumerate(want))


@pytest.mark.parametrize("key,name", [("cyclegan_64", "CycleGAN"), ("attentiongan_64_identity", "AttentionGAN")])
def test_train_cycle_matches_reference_golden(key, name):
    """Model.train_cycle (model.py:660-758) through the drop-in modules: initial weights identical to the reference,
    per-step losses of the first steps within the bf16 bound of the unmodified reference's golden losses."""
    import json
    from models import model as M
    from models.data import SyntheticLoader
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))[key]
    m = M.Model(model=name, topography="all", num_epochs=200, seed=47, add_identity_loss=gold["identity"])
    for net_name in ("pre_to_post_generator", "post_to_pre_generator", "pre_discriminator", "post_discriminator"):
        sd = getattr(m, net_name).state_dict()
        assert set(sd) == set(gold["init"][net_name])
        for k, v in sd.items():
            want = gold["init"][net_name][k]
            assert abs(v.double().abs().sum().item() - want["abs_sum"]) <= 1e-6 * max(1.0, want["abs_sum"])
    loader = list(SyntheticLoader(steps=len(gold["losses"]), batch=gold["batch"], size=gold["size"]))
    for step, batch in enumerate(loader):
        m.train_loader = [batch]
        m.starting_epoch = m.num_epochs  # one epoch == one step; lr stays 2e-4 (lambda_rule with num_epochs=200)
        m.all_losses = m.initialise_loss_storage(overall=True)
        m.train_cycle()
        got = [m.all_losses[k][-1] for k in gold["loss_keys"]]
        print(f"\n[parity] {name} step {step}: " + ", ".join(f"{g:.4f}/{w:.4f}" for g, w in zip(got, gold["losses"][step])))
        tol = LOSS_RTOL if step == 0 else 3 * LOSS_RTOL  # later steps are free-running (see the paired test)
        for k, g, w in zip(gold["loss_keys"], got, gold["losses"][step]):
            assert abs(g - w) <= tol * abs(w) + 1e-3, f"{name} step {step} {k}: {g} vs reference {w}"


def test_paired_100_steps_tracked_against_oracle():
    """north_star: "bf16 rtol 2e-2, tracked over 100 steps". GAN training is chaotic (Adam's updates are ~lr*sign(g)),
    so two runs are reported: (a) TEACHER-FORCED -- before every step the native trainer adopts the oracle's weights
    and Adam moments, so each of the 100 steps is an independent parity check of one full train_paired iteration
    (asserted: every loss within rtol 2e-2, output within OUT_TOL); (b) FREE-RUNNING -- 100 native steps from the
    common initial weights with no resynchronisation (asserted: first step within rtol 2e-2, and the mean of each
    loss over the 100 steps within 5% of the oracle's mean -- the trajectories stay statistically together)."""
    from fpgan.trainer import PairedTrainer
    steps, batch, size = 100, 2, 64
    keys = PairedTrainer.LOSS_KEYS
    O, nets, G, D = make_pair()
    otr = O.PairedTrainer(nets)
    tr = PairedTrainer(G, D)
    O2, nets2, G2, D2 = make_pair()
    free = PairedTrainer(G2, D2)

    def adopt(fp, params, adam):
        """copy oracle parameters + Adam state into the native flat buffers"""
        plist = [v for v in params.values() if v.is_floating_point() and v.dim() > 0]
        for (name, p), src, m, v in zip(fp.named, plist, adam.m, adam.v):
            off, k = fp.offsets[name]
            assert p.shape == src.shape, name
            fp.flat[off:off + k].copy_(src.reshape(-1))
            fp.m[off:off + k].copy_(m.reshape(-1))
            fp.v[off:off + k].copy_(v.reshape(-1))
        fp.steps = adam.t

    worst_loss, worst_out = 0.0, 0.0
    ref_hist, free_hist = [], []
    for step in range(steps):
        x, y = O.synthetic_batch(step, batch, 9, size)
        adopt(tr.gp, otr.G, otr.opt_g)
        adopt(tr.dp, otr.D, otr.opt_d)
        tr._force_repack(tr.G)
        tr._force_repack(tr.D)
        ref = otr.step(x, y)
        synth = tr.step(x.cuda(), y.cuda())
        got = tr.losses()
        e_out = rel_rms(synth, ref["synthetic"])
        worst_out = max(worst_out, e_out)
        for k in keys:
            rel = abs(got[k] - ref[k]) / (abs(ref[k]) + 1e-6)
            worst_loss = max(worst_loss, rel)
            assert rel <= LOSS_RTOL + 1e-4, f"teacher-forced step {step} {k}: {got[k]} vs oracle {ref[k]}"
        assert e_out < OUT_TOL, f"teacher-forced step {step}: output rel-rms {e_out}"
        free.step(x.cuda(), y.cuda())
        fl = free.losses()
        if step == 0:
            for k in keys:
                assert abs(fl[k] - ref[k]) <= LOSS_RTOL * abs(ref[k]) + 1e-4
        ref_hist.append([ref[k] for k in keys])
        free_hist.append([fl[k] for k in keys])
    ref_mean = torch.tensor(ref_hist).mean(0)
    free_mean = torch.tensor(free_hist).mean(0)
    print(f"\n[parity] 100 teacher-forced steps: worst loss rel err {worst_loss:.4f}, worst output rel-rms {worst_out:.4f}")
    print("[parity] 100 free-running steps: mean losses native " + ", ".join(f"{v:.4f}" for v in free_mean.tolist()) +
          " | oracle " + ", ".join(f"{v:.4f}" for v in ref_mean.tolist()))
    for k, a, b in zip(keys, free_mean.tolist(), ref_mean.tolist()):
        assert abs(a - b) <= 0.05 * abs(b) + 1e-3, f"free-running mean of {k}: {a} vs oracle {b}"


# ------------------------------------------------------------------------------------------------ Pix2Pix
def _pix2pix_pair(seed=47):
    from oracle import gan_oracle as O
    from models import model_architectures as A
    nets = O.init_model("pix2pix", "all", seed=seed)
    G, D = A.Pix2PixGenerator(9), A.Pix2PixDiscriminator(9)
    G.load_state_dict(nets["generator"])
    D.load_state_dict(nets["discriminator"])
    return O, nets, G.cuda(), D.cuda()


def _dropout_masks(batch, seed):
    """keep masks of the three dropout blocks in execution order (levels 6, 5, 4: 4x4, 8x8, 16x16 at 256x256)"""
    g = torch.Generator().manual_seed(seed)
    keeps = [torch.rand(batch, 512, s, s, generator=g) < 0.5 for s in (4, 8, 16)]
    oracle = [k.float() * 2.0 for k in keeps]
    native = [k.permute(0, 2, 3, 1).contiguous().to(torch.uint8).reshape(-1).cuda() for k in keeps]
    return oracle, native


def test_pix2pix_constructors_match_reference_init():
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))["pix2pix_256"]
    from models import model_architectures as A
    from models.model import Model
    torch.manual_seed(47)
    G = A.Pix2PixGenerator(9).apply(Model.initialise_weights)
    D = A.Pix2PixDiscriminator(9).apply(Model.initialise_weights)
    for net, key in ((G, "generator"), (D, "discriminator")):
        sd = net.state_dict()
        assert set(k for k, v in sd.items() if v.is_floating_point()) == set(gold["init"][key])
        for k, want in gold["init"][key].items():
            assert abs(sd[k].double().sum().item() - want["sum"]) <= 1e-9 * max(1.0, abs(want["sum"])), k


def test_pix2pix_forward_and_step_match_oracle():
    """BASELINE.json configs[0]: Pix2Pix U-Net + BatchNorm PatchGAN, batch 1, 256x256, topography=all -- forward
    tensors, BatchNorm running statistics and the losses of a full train_paired step (module / autograd path)
    against the oracle with identical dropout masks."""
    O, nets, G, D = _pix2pix_pair()
    x, y = O.synthetic_batch(0, 1, 9, 256)
    om, nm = _dropout_masks(1, 5)
    G._executor().dropout_masks = nm
    ref_nets = {k: {n: v.clone() for n, v in p.items()} for k, p in nets.items()}
    with torch.no_grad():
        ref_out = O.pix2pix_generator_forward(ref_nets["generator"], x, masks=om)
        ref_logits = O.patchgan_bn_forward(ref_nets["discriminator"], torch.cat((x, y), 1))
        out = G(x.cuda())
        logits = D(torch.cat((x, y), 1).cuda())
    e_out, e_log = rel_rms(out, ref_out), rel_rms(logits, ref_logits)
    print(f"[parity] pix2pix generator rel-rms {e_out:.4f}, BatchNorm PatchGAN logits rel-rms {e_log:.4f}")
    assert e_out < 6e-2 and e_log < OUT_TOL
    for net, ref in ((G, ref_nets["generator"]), (D, ref_nets["discriminator"])):
        sd = net.state_dict()
        for k in ref:
            if "running_" in k:
                assert rel_rms(sd[k], ref[k]) < 2e-2, k
            if "num_batches" in k:
                assert int(sd[k]) == int(ref[k]) == 1

    # one full training step through the module path (losses + updated parameters)
    O, nets, G, D = _pix2pix_pair()
    G._executor().dropout_masks = nm
    tr = O.PairedTrainer(nets, "pix2pix")
    tr.g_forward = lambda p, inp: O.pix2pix_generator_forward(p, inp, masks=om)
    ref = tr.step(x, y)
    opt_d = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_g = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    xc, yc = x.cuda(), y.cuda()
    mse = torch.nn.MSELoss()
    syn = G(xc)
    d_syn_in, d_real_in = torch.cat((xc, syn), 1), torch.cat((xc, yc), 1)
    opt_d.zero_grad()
    p_s = D(d_syn_in.detach())
    l_syn = mse(p_s, torch.zeros_like(p_s))
    p_r = D(d_real_in)
    l_real = mse(p_r, torch.ones_like(p_r))
    ((l_syn + l_real) * 0.5).backward()
    opt_d.step()
    for p in D.parameters():
        p.requires_grad = False
    opt_g.zero_grad()
    p_g = D(d_syn_in)
    l_adv = mse(p_g, torch.ones_like(p_g))
    l_l1 = torch.nn.functional.l1_loss(syn, yc) * 100
    (l_adv + l_l1).backward()
    opt_g.step()
    got = {"losses_discriminator_real": l_real.item(), "losses_discriminator_synthetic": l_syn.item(),
           "losses_generator_synthetic": l_adv.item(), "l1_losses_generator_synthetic": l_l1.item()}
    print("[parity] pix2pix step losses", got, "oracle", {k: ref[k] for k in got})
    for k in got:
        assert abs(got[k] - ref[k]) <= 3e-2 * abs(ref[k]) + 1e-3, (k, got[k], ref[k])
    # Adam's first step moves every weight by ~lr regardless of the gradient scale: compare the update DIRECTION
    agree = []
    for (n_, p), q in zip(G.named_parameters(), [v for k, v in tr.G.items() if ".running_" not in k
                                                  and v.is_floating_point() and v.dim() > 0]):
        init = O.init_model("pix2pix", "all", seed=47)["generator"][n_]
        d_nat, d_ref = (p.detach().cpu() - init).flatten(), (q.detach() - init).flatten()
        agree.append(((d_nat * d_ref) > 0).float().mean().item())
    print(f"[parity] pix2pix Adam update sign agreement: min {min(agree):.3f} mean {sum(agree) / len(agree):.3f}")
    assert sum(agree) / len(agree) > 0.8


def test_fused_pix2pix_step_matches_oracle_teacher_forced():
    """Pix2PixTrainer (fused train_paired for the BatchNorm / dropout model, model.py:611-651) against the oracle's
    fp32 step with identical dropout masks, TEACHER-FORCED: before every step the native trainer adopts the oracle's
    weights, Adam moments and BatchNorm running statistics, so each of the five steps -- eager 0-1, the captured step 2,
    graph replays afterwards -- checks one full iteration: the four losses (rtol 3e-2: BatchNorm over 4-64 values per
    channel at the bottleneck with B=1 amplifies bf16 noise), the generated image, and the
    BatchNorm running statistics after the step (three discriminator calls, one generator call)."""
    from fpgan.trainer import Pix2PixTrainer
    O, nets, G, D = _pix2pix_pair()
    otr = O.PairedTrainer(nets, "pix2pix")
    tr = Pix2PixTrainer(G, D)

    def adopt(fp, module, params, adam):
        plist = [v for k_, v in params.items() if v.is_floating_point() and v.dim() > 0 and "running_" not in k_]
        assert len(plist) == len(fp.named)
        for (name, p), src, m, v in zip(fp.named, plist, adam.m, adam.v):
            off, k = fp.offsets[name]
            assert p.shape == src.shape, name
            fp.flat[off:off + k].copy_(src.detach().reshape(-1))
            fp.m[off:off + k].copy_(m.reshape(-1))
            fp.v[off:off + k].copy_(v.reshape(-1))
        fp.steps = adam.t
        sd = module.state_dict()
        for k_, v_ in params.items():
            if "running_" in k_ or "num_batches" in k_:
                sd[k_].copy_(v_)

    worst = 0.0
    for step in range(5):
        adopt(tr.gp, G, otr.G, otr.opt_g)
        adopt(tr.dp, D, otr.D, otr.opt_d)
        tr.G.repack(force=True)
        tr.D.repack(force=True)
        x, y = O.synthetic_batch(step, 1, 9, 256)
        om, nm = _dropout_masks(1, 5)  # the same masks every step: under graph replay the injected buffers are baked in
        if tr.inject_masks is None:
            tr.inject_masks = nm
        otr.g_forward = lambda p, inp: O.pix2pix_generator_forward(p, inp, masks=om)
        ref = otr.step(x, y)
        out = tr.step(x.cuda(), y.cuda())
        got = tr.losses()
        e = rel_rms(out, ref["synthetic"])
        print(f"\n[parity] pix2pix fused step {step}: synthetic rel-rms err {e:.4f}; losses " +
              ", ".join(f"{got[k]:.5f}/{ref[k]:.5f}" for k in tr.LOSS_KEYS))
        assert e < 6e-2, e
        for k in tr.LOSS_KEYS:
            worst = max(worst, abs(got[k] - ref[k]) / (abs(ref[k]) + 1e-6))
            assert abs(got[k] - ref[k]) <= 3e-2 * abs(ref[k]) + 1e-3, f"step {step} {k}: {got[k]} vs {ref[k]}"
        for module, params in ((G, otr.G), (D, otr.D)):
            sd = module.state_dict()
            for k_, v_ in params.items():
                if "running_" in k_:
                    assert rel_rms(sd[k_], v_) < 2e-2, (step, k_)
                if "num_batches" in k_:
                    assert int(sd[k_]) == int(v_), (step, k_)
    assert tr._graphs, "the step was never captured"
    print(f"[parity] pix2pix fused: worst loss rel err over 5 teacher-forced steps {worst:.5f}")


def test_fused_pix2pix_draws_new_dropout_masks_on_every_replay():
    """the captured step reads its dropout seeds from device memory: two replays on the same batch from the same
    weights give different images (fresh masks), the same torch seed gives the same image"""
    from fpgan.trainer import Pix2PixTrainer
    O, nets, G, D = _pix2pix_pair()
    tr = Pix2PixTrainer(G, D)
    x, y = (t.cuda() for t in O.synthetic_batch(0, 1, 9, 256))
    outs = []
    for step in range(5):
        torch.manual_seed(100 + (step if step < 4 else 3))  # steps 3 and 4 share their seed
        outs.append(tr.step(x, y, lr_g=0.0, lr_d=0.0).clone())  # lr 0: the weights stay what they are
    assert tr._graphs
    assert not torch.equal(outs[2], outs[3])
    assert torch.equal(outs[3], outs[4])


def test_model_api_pix2pix_train_paired_and_checkpoint(tmp_path):
    """`--model=Pix2Pix` through the public API: Model(...).train_paired() on synthetic 256x256 batches, losses stay
    in the reference's range (golden first-step losses of the unmodified reference; dropout masks differ: the
    reference draws them from the CPU RNG), BatchNorm buffers land in the checkpoint."""
    import json
    from models import model as M
    from models.data import SyntheticLoader
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))["pix2pix_256"]
    m = M.Model(model="Pix2Pix", topography="all", num_epochs=200, seed=47, data_path=str(tmp_path),
                save_model_interval=200)
    m.train_loader = list(SyntheticLoader(steps=2, batch=1, size=256))
    m.starting_epoch = m.num_epochs
    m.train_paired()
    got = [m.all_losses[k][-1] for k in ("all_losses_discriminator_real", "all_losses_discriminator_synthetic",
                                         "all_losses_generator_synthetic", "all_l1_losses_generator_synthetic")]
    want = [sum(s[i] for s in gold["losses"]) / len(gold["losses"]) for i in range(4)]
    print("[parity] Pix2Pix epoch-mean losses", got, "reference (other dropout masks)", want)
    for g, w in zip(got, want):
        assert abs(g - w) <= 0.1 * abs(w), (got, want)
    files = list((tmp_path / "models").glob("Pix2Pix_*.pth.tar"))
    assert len(files) == 1
    ck = torch.load(files[0], weights_only=False)
    assert len(ck["generator"]) == 82 and int(ck["generator"]["model.model.1.model.2.num_batches_tracked"]) == 2
    assert int(ck["discriminator"]["model.3.num_batches_tracked"]) == 6  # three discriminator calls per step


# ------------------------------------------------------------------------------------------------ segmentation U-Net
def test_pix2pix_dropout_masks_follow_the_torch_seed():
    """The reference seeds torch with 47 before every evaluation-time generator call (model.py:393,497,579) because
    Pix2Pix's dropout stays active: the same seed must give the same image whatever ran before, another seed another."""
    from models import model_architectures as A
    torch.manual_seed(3)
    G = A.Pix2PixGenerator(9).cuda()
    x = torch.rand(1, 9, 256, 256, generator=torch.Generator().manual_seed(5)).cuda() * 2 - 1
    with torch.no_grad():
        torch.manual_seed(47)
        a = G(x).clone()
        G(x)  # an unrelated forward in between advances the generator
        torch.manual_seed(47)
        b = G(x).clone()
        torch.manual_seed(48)
        c = G(x).clone()
    assert torch.equal(a, b)
    assert not torch.equal(a, c)


def test_unet_inference_masks_and_counts():
    """BASELINE.json configs[4] at test size: the segmentation U-Net of calculate_metrics (model.py:380-418) on a
    generated / ground-truth pair. Logits against the oracle within the bf16 bound; the integer work -- the
    (sigmoid > 0.5) threshold and the TP/FP/TN/FN counts -- bit-exact against the reference expressions evaluated on
    the same logits."""
    from oracle import gan_oracle as O
    from models import model_architectures as A
    p = O.init_unet(47)
    # make the head decisive: with N(0, 0.02) weights the logits sit within bf16 noise of the threshold
    p["outc.conv.weight"] = p["outc.conv.weight"] * 40
    net = A.UNet()
    net.load_state_dict(p)
    net = net.cuda()
    g = torch.Generator().manual_seed(2000)
    gen = torch.rand(4, 3, 64, 64, generator=g) * 2 - 1
    truth = torch.rand(4, 3, 64, 64, generator=g) * 2 - 1
    ref = {k: v.clone() for k, v in p.items()}
    with torch.no_grad():
        ref_logits = O.unet_forward(ref, torch.clamp((gen + 1) * 0.5, min=0, max=1))
    logits = net(torch.clamp((gen.cuda() + 1) * 0.5, min=0, max=1))
    err = rel_rms(logits - logits.mean(), ref_logits - ref_logits.mean())
    print(f"[parity] U-Net logits rel-rms {err:.4f} (spread {ref_logits.std().item():.3f})")
    assert err < 5e-2
    sd = net.state_dict()
    assert int(sd["inc.double_conv.1.num_batches_tracked"]) == 1
    assert rel_rms(sd["down4.maxpool_conv.1.double_conv.4.running_var"],
                   ref["down4.maxpool_conv.1.double_conv.4.running_var"]) < 3e-2
    # away from the threshold the masks agree with the oracle's
    mo, mt, counts = A.flood_masks_and_counts(net, gen.cuda(), truth.cuda())
    safe = ref_logits.abs() > 6 * (logits.cpu() - ref_logits).abs().max()
    assert safe.float().mean() > 0.3
    assert torch.equal(mo.cpu()[safe], O.flood_mask(ref_logits)[safe])
    # bit-exact integer work on the native logits (second forward pass: BatchNorm buffers moved, logits unchanged
    # because batch statistics are used)
    lg = net(torch.clamp((gen.cuda() + 1) * 0.5, min=0, max=1))
    lt = net(torch.clamp((truth.cuda() + 1) * 0.5, min=0, max=1))
    assert torch.equal(mo, (torch.sigmoid(lg) > 0.5).float()) and torch.equal(mt, (torch.sigmoid(lt) > 0.5).float())
    assert counts.tolist() == O.confusion_counts(mo.flatten().cpu(), mt.flatten().cpu())


def test_unet_masks_against_the_fp32_oracle_without_tricks():
    """The same comparison with the head as initialised (no x40): the logits of a randomly initialised U-Net sit close
    to the threshold, so some pixels legitimately flip under bf16 noise. Reported: the mask disagreement rate and the
    confusion-count deltas against the fp32 oracle. Asserted: every disagreeing pixel lies within the logit error band
    (|fp32 logit| <= max |logit error|) -- i.e. the masks differ only where bf16 cannot decide -- and the rate is what
    that band predicts."""
    from oracle import gan_oracle as O
    from models import model_architectures as A
    p = O.init_unet(47)
    net = A.UNet()
    net.load_state_dict(p)
    net = net.cuda()
    g = torch.Generator().manual_seed(2001)
    gen = torch.rand(4, 3, 128, 128, generator=g) * 2 - 1
    truth = torch.rand(4, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        ref_masks = O.segmentation_masks({k: v.clone() for k, v in p.items()}, gen, truth)
        ref_logits = O.unet_forward({k: v.clone() for k, v in p.items()}, torch.clamp((gen + 1) * 0.5, min=0, max=1))
    logits = net(torch.clamp((gen.cuda() + 1) * 0.5, min=0, max=1)).cpu()
    mo, mt, counts = A.flood_masks_and_counts(net, gen.cuda(), truth.cuda())
    band = (logits - ref_logits).abs().max().item()
    differ = mo.cpu() != ref_masks[0]
    rate = differ.float().mean().item()
    in_band = (ref_logits.abs() <= band).float().mean().item()
    ref_counts = O.confusion_counts(ref_masks[0].flatten(), ref_masks[1].flatten())
    print(f"\n[parity] U-Net masks vs fp32 oracle (head as initialised): disagreement {rate:.4%} of pixels, "
          f"{in_band:.4%} of fp32 logits lie within the error band +-{band:.2e}; confusion counts native "
          f"{counts.tolist()} oracle {ref_counts}")
    assert (ref_logits[differ].abs() <= band).all(), "a mask pixel flipped although its fp32 logit is decisive"
    assert rate <= in_band


def test_model_calculate_metrics_flood_columns():
    """evaluate.py --calculate_metrics path (model.py:363-422): generator inference, segmentation of generated and
    ground-truth tiles, bit-exact masks, confusion counts over the split, torchmetrics-equivalent flood metrics."""
    from models import model as M
    from models.data import SyntheticLoader
    m = M.Model(model="PairedAttention", topography="all", num_epochs=1, seed=47, training_model=False)
    res = m.calculate_metrics(loader=list(SyntheticLoader(steps=2, batch=2, size=64)))
    for k in ("MSE", "Accuracy", "F1_Flood", "Precision_Flood", "Recall_Flood", "F1_No_Flood", "Precision_No_Flood",
              "Recall_No_Flood", "Inference"):
        assert 0.0 <= res[k] <= 1.0 or k == "Inference", (k, res[k])
    assert abs(res["MSE"] + res["Accuracy"] - 1.0) < 1e-12  # both are functions of the same integer counts
    assert 0.0 < res["PSNR"] < 60.0 and -1.0 <= res["SSIM"] <= 1.0  # image-quality columns (models/metrics.py)
    assert res["MS-SSIM"] != res["MS-SSIM"] and res["LPIPS"] != res["LPIPS"]  # 64-pixel tiles are below MS-SSIM's 160
