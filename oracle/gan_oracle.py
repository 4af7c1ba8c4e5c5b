"""ORACLE -- test infrastructure only. Never imported by the product path.

A CPU restatement, in plain fp32 PyTorch functional ops, of the GAN training-step path of
Natasha-R/Flood-Prediction-GAN (models/model_architectures.py, models/model.py). Every function cites the
reference lines it follows. Parameters live in ordinary dicts keyed by the reference's state_dict names, so a
reference checkpoint loads directly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Parity is PINNED: tests/golden/*.json hold outputs of the unmodified reference (imported from /root/reference by
tests/golden/make_golden.py in the authoring container) and tests/test_oracle_cpu.py checks this restatement
against them (and against the live reference when /root/reference is present).
"""
import math
import random
from collections import OrderedDict

import torch
import torch.nn.functional as F

TOPOGRAPHY_CHANNELS = {"all": 9, "map": 6, "dem": 4, "flow": 4, "river": 4, None: 3}  # model.py:78


# ---------------------------------------------------------------------------------------------- initialisation
def _conv_ctor_draw(shape, bias):
    """RNG consumption of nn.Conv2d / nn.ConvTranspose2d.__init__ (reset_parameters: kaiming_uniform_ on the
    weight, uniform_ on the bias). The values are overwritten by initialise_weights; only the draws matter."""
    torch.empty(shape).uniform_(-1, 1)
    if bias:
        torch.empty(shape[0] if len(shape) == 4 else 1).uniform_(-1, 1)


def _normal(shape, mean=0.0):
    return torch.empty(shape).normal_(mean, 0.02)  # model.py:168 / :172


def _init_from_plan(plan):
    """plan: list of (name, weight_shape, has_bias, kind) in CONSTRUCTION order, kind in {conv, convT, bn}.
    Returns the parameters after `.apply(initialise_weights)` (model.py:162-173) which visits the same order."""
    for name, shape, bias, kind in plan:
        if kind == "bn":
            continue  # BatchNorm2d.__init__ draws nothing
        if kind == "convT":
            # bias of ConvTranspose2d has out_channels = shape[1] elements
            torch.empty(shape).uniform_(-1, 1)
            if bias:
                torch.empty(shape[1]).uniform_(-1, 1)
        else:
            _conv_ctor_draw(shape, bias)
    params = OrderedDict()
    for name, shape, bias, kind in plan:
        if kind == "bn":
            params[name + ".weight"] = _normal(shape, 1.0)
            params[name + ".bias"] = torch.zeros(shape)
            params[name + ".running_mean"] = torch.zeros(shape)
            params[name + ".running_var"] = torch.ones(shape)
            params[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        else:
            params[name + ".weight"] = _normal(shape)
            if bias:
                params[name + ".bias"] = torch.zeros(shape[1] if kind == "convT" else shape[0])
    return params


def attention_generator_plan(input_channels):
    """PairedAttentionGenerator / AttentionGANGenerator (identical code): model_architectures.py:312-334, 170-192"""
    p = [("conv1", (64, input_channels, 7, 7), True, "conv"), ("conv2", (128, 64, 3, 3), True, "conv"),
         ("conv3", (256, 128, 3, 3), True, "conv")]
    for i in range(9):
        p += [(f"resnet_blocks.{i}.conv1", (256, 256, 3, 3), True, "conv"),
              (f"resnet_blocks.{i}.conv2", (256, 256, 3, 3), True, "conv")]
    p += [("deconv1_content", (256, 128, 3, 3), True, "convT"), ("deconv2_content", (128, 64, 3, 3), True, "convT"),
          ("deconv3_content", (27, 64, 7, 7), True, "conv"),
          ("deconv1_attention", (256, 128, 3, 3), True, "convT"), ("deconv2_attention", (128, 64, 3, 3), True, "convT"),
          ("deconv3_attention", (10, 64, 1, 1), True, "conv")]
    return p


def cyclegan_generator_plan(input_channels):
    """CycleGANGenerator: model_architectures.py:95-117 (Sequential indices are the state_dict keys)"""
    p = [("model.1", (64, input_channels, 7, 7), True, "conv"), ("model.4", (128, 64, 3, 3), True, "conv"),
         ("model.7", (256, 128, 3, 3), True, "conv")]
    for i in range(9):
        p += [(f"model.{10 + i}.conv_block.1", (256, 256, 3, 3), True, "conv"),
              (f"model.{10 + i}.conv_block.5", (256, 256, 3, 3), True, "conv")]
    p += [("model.19", (256, 128, 3, 3), True, "convT"), ("model.22", (128, 64, 3, 3), True, "convT"),
          ("model.26", (3, 64, 7, 7), True, "conv")]
    return p


def patchgan_plan(in_channels, batch_norm):
    """70x70 PatchGAN: model_architectures.py:68-81 (Pix2Pix: BatchNorm, no bias on normed convs), :140-153, :282-295,
    :424-437 (InstanceNorm, bias everywhere)."""
    p = [("model.0", (64, in_channels, 4, 4), True, "conv"),
         ("model.2", (128, 64, 4, 4), not batch_norm, "conv")]
    if batch_norm:
        p.append(("model.3", (128,), False, "bn"))
    p.append(("model.5", (256, 128, 4, 4), not batch_norm, "conv"))
    if batch_norm:
        p.append(("model.6", (256,), False, "bn"))
    p.append(("model.8", (512, 256, 4, 4), not batch_norm, "conv"))
    if batch_norm:
        p.append(("model.9", (512,), False, "bn"))
    p.append(("model.11", (1, 512, 4, 4), True, "conv"))
    return p


# Pix2Pix U-Net (model_architectures.py:9-63). Blocks from the outside in: (outer_nc, inner_nc, kind)
PIX2PIX_BLOCKS = [(3, 64, "outermost"), (64, 128, "middle"), (128, 256, "middle"), (256, 512, "middle"),
                  (512, 512, "dropout"), (512, 512, "dropout"), (512, 512, "dropout"), (512, 512, "innermost")]


def pix2pix_block_prefix(level):
    """state_dict prefix of block `level` (0 = outermost): the sub-block is child 1 of the outermost Sequential and
    child 3 of every other one (model_architectures.py:41,53-55)."""
    prefix = "model.model."
    for lv in range(level):
        prefix += ("1." if lv == 0 else "3.") + "model."
    return prefix


def pix2pix_layer_names(level):
    """(downconv, downnorm, upconv, upnorm) state_dict names of block `level` (None where the layer does not exist)."""
    pre = pix2pix_block_prefix(level)
    kind = PIX2PIX_BLOCKS[level][2]
    if kind == "outermost":
        return pre + "0", None, pre + "3", None           # [downconv, sub, uprelu, upconv, tanh]
    if kind == "innermost":
        return pre + "1", None, pre + "3", pre + "4"      # [downrelu, downconv, uprelu, upconv, upnorm]
    return pre + "1", pre + "2", pre + "5", pre + "6"     # [downrelu, downconv, downnorm, sub, uprelu, upconv, upnorm, (dropout)]


def _bn_entries(params, name, c):
    params[name + ".weight"] = _normal((c,), 1.0)
    params[name + ".bias"] = torch.zeros(c)
    params[name + ".running_mean"] = torch.zeros(c)
    params[name + ".running_var"] = torch.ones(c)
    params[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def init_pix2pix_generator(input_channels):
    """Pix2PixGenerator.__init__ (model_architectures.py:13-19) builds the blocks from the INSIDE out, each drawing
    downconv then upconv (:32,41,46,51); `.apply(initialise_weights)` (model.py:103) then visits the Sequential children
    in registration order: down layers on the way in, up layers on the way out."""
    shapes = {}
    for level in reversed(range(8)):
        outer, inner, kind = PIX2PIX_BLOCKS[level]
        in_nc = input_channels if kind == "outermost" else outer
        down = (inner, in_nc, 4, 4)
        up = (inner if kind == "innermost" else inner * 2, outer, 4, 4)  # ConvTranspose2d weight [Cin][Cout][4][4]
        shapes[level] = (down, up)
        torch.empty(down).uniform_(-1, 1)
        torch.empty(up).uniform_(-1, 1)
        if kind == "outermost":
            torch.empty(outer).uniform_(-1, 1)  # the only layer with a bias (:41)
    draws = {}

    def visit(level):  # nn.Module.apply order
        outer, inner, kind = PIX2PIX_BLOCKS[level]
        dn, dnorm, un, unorm = pix2pix_layer_names(level)
        draws[dn + ".weight"] = _normal(shapes[level][0])
        if dnorm:
            _bn_entries(draws, dnorm, inner)
        if kind != "innermost":
            visit(level + 1)
        draws[un + ".weight"] = _normal(shapes[level][1])
        if kind == "outermost":
            draws[un + ".bias"] = torch.zeros(outer)
        if unorm:
            _bn_entries(draws, unorm, outer)

    visit(0)
    params = OrderedDict()

    def order(level):  # state_dict order: registration order, depth first
        outer, inner, kind = PIX2PIX_BLOCKS[level]
        dn, dnorm, un, unorm = pix2pix_layer_names(level)
        keys = [dn + ".weight"]
        if dnorm:
            keys += [dnorm + s for s in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")]
        for k in keys:
            params[k] = draws[k]
        if kind != "innermost":
            order(level + 1)
        keys = [un + ".weight"] + ([un + ".bias"] if kind == "outermost" else [])
        if unorm:
            keys += [unorm + s for s in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")]
        for k in keys:
            params[k] = draws[k]

    order(0)
    return params


def init_model(model, topography="all", seed=47, training=True):
    """Model.__init__ wiring (model.py:78-104): seed once, then generator(s) and discriminator(s) in this order."""
    n = TOPOGRAPHY_CHANNELS[topography]
    torch.manual_seed(seed)
    nets = OrderedDict()
    if model == "pairedattention":
        nets["generator"] = _init_from_plan(attention_generator_plan(n))
        if training:
            nets["discriminator"] = _init_from_plan(patchgan_plan(n + 3, False))
    elif model in ("attentiongan", "cyclegan"):
        plan = attention_generator_plan if model == "attentiongan" else cyclegan_generator_plan
        nets["pre_to_post_generator"] = _init_from_plan(plan(n))
        nets["post_to_pre_generator"] = _init_from_plan(plan(n))
        if training:
            nets["pre_discriminator"] = _init_from_plan(patchgan_plan(n, False))
            nets["post_discriminator"] = _init_from_plan(patchgan_plan(n, False))
    elif model == "pix2pix":
        nets["generator"] = init_pix2pix_generator(n)
        if training:
            nets["discriminator"] = _init_from_plan(patchgan_plan(n + 3, True))
    else:
        raise NotImplementedError(model)
    return nets


# ---------------------------------------------------------------------------------------------- forward passes
def _in(x):
    return F.instance_norm(x, eps=1e-5)  # nn.InstanceNorm2d defaults: no affine, no running stats


def attention_generator_forward(p, x, return_mask=False):
    """model_architectures.py:339-400 (== :197-258)"""
    inp = x
    x = F.pad(x, (3, 3, 3, 3), "reflect")
    x = F.relu(_in(F.conv2d(x, p["conv1.weight"], p["conv1.bias"])))
    x = F.relu(_in(F.conv2d(x, p["conv2.weight"], p["conv2.bias"], stride=2, padding=1)))
    x = F.relu(_in(F.conv2d(x, p["conv3.weight"], p["conv3.bias"], stride=2, padding=1)))
    for i in range(9):  # PairedAttentionBlock.forward :412-418
        b = f"resnet_blocks.{i}."
        y = F.pad(x, (1, 1, 1, 1), "reflect")
        y = F.relu(_in(F.conv2d(y, p[b + "conv1.weight"], p[b + "conv1.bias"])))
        y = F.pad(y, (1, 1, 1, 1), "reflect")
        y = _in(F.conv2d(y, p[b + "conv2.weight"], p[b + "conv2.bias"]))
        x = x + y

    def up(t, name):
        return F.relu(_in(F.conv_transpose2d(t, p[name + ".weight"], p[name + ".bias"], stride=2, padding=1,
                                             output_padding=1)))
    c = up(up(x, "deconv1_content"), "deconv2_content")
    c = F.pad(c, (3, 3, 3, 3), "reflect")
    content = torch.tanh(F.conv2d(c, p["deconv3_content.weight"], p["deconv3_content.bias"]))
    a = up(up(x, "deconv1_attention"), "deconv2_attention")
    att = torch.softmax(F.conv2d(a, p["deconv3_attention.weight"], p["deconv3_attention.bias"]), dim=1)
    # :383-399 -- sum in the reference's order: output1 + ... + output9 + output10
    out = content[:, 0:3] * att[:, 0:1]
    for k in range(1, 9):
        out = out + content[:, 3 * k:3 * k + 3] * att[:, k:k + 1]
    out = out + inp[:, :3] * att[:, 9:10]
    if return_mask:
        return out, att[:, 9]
    return out


def cyclegan_generator_forward(p, x):
    """model_architectures.py:95-134"""
    x = F.pad(x, (3, 3, 3, 3), "reflect")
    x = F.relu(_in(F.conv2d(x, p["model.1.weight"], p["model.1.bias"])))
    x = F.relu(_in(F.conv2d(x, p["model.4.weight"], p["model.4.bias"], stride=2, padding=1)))
    x = F.relu(_in(F.conv2d(x, p["model.7.weight"], p["model.7.bias"], stride=2, padding=1)))
    for i in range(9):
        b = f"model.{10 + i}.conv_block."
        y = F.pad(x, (1, 1, 1, 1), "reflect")
        y = F.relu(_in(F.conv2d(y, p[b + "1.weight"], p[b + "1.bias"])))
        y = F.pad(y, (1, 1, 1, 1), "reflect")
        y = _in(F.conv2d(y, p[b + "5.weight"], p[b + "5.bias"]))
        x = x + y
    for name in ("model.19", "model.22"):
        x = F.relu(_in(F.conv_transpose2d(x, p[name + ".weight"], p[name + ".bias"], stride=2, padding=1,
                                          output_padding=1)))
    x = F.pad(x, (3, 3, 3, 3), "reflect")
    return torch.tanh(F.conv2d(x, p["model.26.weight"], p["model.26.bias"]))


def patchgan_forward(p, x):
    """InstanceNorm PatchGAN: model_architectures.py:424-441 (== :140-157, :282-299)"""
    x = F.leaky_relu(F.conv2d(x, p["model.0.weight"], p["model.0.bias"], stride=2, padding=1), 0.2)
    x = F.leaky_relu(_in(F.conv2d(x, p["model.2.weight"], p["model.2.bias"], stride=2, padding=1)), 0.2)
    x = F.leaky_relu(_in(F.conv2d(x, p["model.5.weight"], p["model.5.bias"], stride=2, padding=1)), 0.2)
    x = F.leaky_relu(_in(F.conv2d(x, p["model.8.weight"], p["model.8.bias"], stride=1, padding=1)), 0.2)
    return F.conv2d(x, p["model.11.weight"], p["model.11.bias"], stride=1, padding=1)


def _bn(p, name, x):
    """nn.BatchNorm2d in training mode: batch statistics, running statistics updated in place (momentum 0.1)."""
    p[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"], p[name + ".bias"],
                        training=True, momentum=0.1, eps=1e-5)


def pix2pix_generator_forward(p, x, masks=None):
    """Pix2PixGenerator (model_architectures.py:9-63) in training mode. The LeakyReLU / ReLU layers are IN PLACE
    (:33-34): downrelu overwrites the block input that torch.cat((x, model(x))) (:63) then reads, and uprelu
    overwrites the concatenated output of the sub-block, so a block returns cat(lrelu(x), up(...)) and the up
    convolution sees relu(cat(lrelu(e), d)) = cat(relu(e), relu(d)).
    masks: optional dropout masks (values 0 or 2) for the three dropout blocks in execution order (levels 6, 5, 4);
    None draws them with F.dropout from the global RNG as the reference does."""
    enc = []  # e_k: output of block k's downconv (+ downnorm)
    t = x
    for level in range(8):
        dn, dnorm, _, _ = pix2pix_layer_names(level)
        if level > 0:
            t = F.leaky_relu(t, 0.2)
        t = F.conv2d(t, p[dn + ".weight"], None, stride=2, padding=1)
        if dnorm:
            t = _bn(p, dnorm, t)
        enc.append(t)
    mi = 0
    d = None
    for level in reversed(range(8)):
        _, _, un, unorm = pix2pix_layer_names(level)
        kind = PIX2PIX_BLOCKS[level][2]
        src = F.relu(enc[level]) if kind == "innermost" else torch.cat((F.relu(enc[level]), F.relu(d)), 1)
        d = F.conv_transpose2d(src, p[un + ".weight"], p.get(un + ".bias"), stride=2, padding=1)
        if unorm:
            d = _bn(p, unorm, d)
        if kind == "dropout":
            if masks is None:
                d = F.dropout(d, 0.5, training=True)
            else:
                d = d * masks[mi]
                mi += 1
    return torch.tanh(d)


def patchgan_bn_forward(p, x):
    """Pix2PixDiscriminator (model_architectures.py:65-85): BatchNorm PatchGAN, no bias on the normalised convs"""
    x = F.leaky_relu(F.conv2d(x, p["model.0.weight"], p["model.0.bias"], stride=2, padding=1), 0.2)
    x = F.leaky_relu(_bn(p, "model.3", F.conv2d(x, p["model.2.weight"], None, stride=2, padding=1)), 0.2)
    x = F.leaky_relu(_bn(p, "model.6", F.conv2d(x, p["model.5.weight"], None, stride=2, padding=1)), 0.2)
    x = F.leaky_relu(_bn(p, "model.9", F.conv2d(x, p["model.8.weight"], None, stride=1, padding=1)), 0.2)
    return F.conv2d(x, p["model.11.weight"], p["model.11.bias"], stride=1, padding=1)


GENERATOR_FORWARD = {"pairedattention": attention_generator_forward, "attentiongan": attention_generator_forward,
                     "cyclegan": cyclegan_generator_forward, "pix2pix": pix2pix_generator_forward}
DISCRIMINATOR_FORWARD = {"pix2pix": patchgan_bn_forward}


# ---------------------------------------------------------------------------------------------- optimiser
class Adam:
    """torch.optim.Adam(lr=2e-4, betas=(0.5, 0.999), eps=1e-8), no weight decay / amsgrad (model.py:112-122),
    restated from the documented update rule; lr may be rescaled per epoch by lambda_rule (model.py:175-181)."""

    def __init__(self, params, lr=0.0002, betas=(0.5, 0.999), eps=1e-8):
        self.params = list(params)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    @torch.no_grad()
    def step(self, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if g is None:
                continue
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / bc1)


def lambda_rule(epoch, num_epochs):
    """model.py:175-181"""
    return 1.0 - max(0, epoch + 1 - (num_epochs / 2)) / float((num_epochs / 2) + 1)


def _float_params(p):
    """trainable tensors: everything but the BatchNorm buffers (running statistics, batch counter)"""
    return [v for k, v in p.items() if v.is_floating_point() and v.dim() > 0 and ".running_" not in k]


def _req(p, flag):
    for v in _float_params(p):
        v.requires_grad_(flag)


# ---------------------------------------------------------------------------------------------- paired step
class PairedTrainer:
    """Model.train_paired inner loop (model.py:611-651): PairedAttention (InstanceNorm) and Pix2Pix (BatchNorm in
    training mode + dropout from the global torch RNG, seeded per epoch at model.py:609)."""

    def __init__(self, nets, model="pairedattention"):
        self.G, self.D = nets["generator"], nets["discriminator"]
        self.g_forward = GENERATOR_FORWARD[model]
        self.d_forward = DISCRIMINATOR_FORWARD.get(model, patchgan_forward)
        self.opt_d = Adam(_float_params(self.D))  # model.py:121
        self.opt_g = Adam(_float_params(self.G))  # model.py:122

    def step(self, input_stack, output_image):
        G, D = self.G, self.D
        _req(G, True)
        synthetic = self.g_forward(G, input_stack)                                   # :615
        concat_real = torch.cat((input_stack, output_image), 1)                      # :616
        concat_synth = torch.cat((input_stack, synthetic), 1)                        # :617
        _req(D, True)                                                                # :620-622
        pred_s = self.d_forward(D, concat_synth.detach())                          # :624
        loss_d_synth = F.mse_loss(pred_s, torch.full(pred_s.shape, 0.0))             # :626-627
        pred_r = self.d_forward(D, concat_real)                                    # :628
        loss_d_real = F.mse_loss(pred_r, torch.full(pred_s.shape, 1.0))              # :629-630
        loss_d = (loss_d_synth + loss_d_real) * 0.5                                  # :631
        dparams = _float_params(D)
        self.opt_d.step(torch.autograd.grad(loss_d, dparams))                        # :632-633
        _req(D, False)                                                               # :636-638
        pred_s = self.d_forward(D, concat_synth)                                   # :640 (updated D)
        loss_g_adv = F.mse_loss(pred_s, torch.full(pred_s.shape, 1.0))               # :641-642
        loss_l1 = F.l1_loss(synthetic, output_image) * 100                           # :643
        gparams = _float_params(G)
        grads = torch.autograd.grad(loss_g_adv + loss_l1, gparams, allow_unused=True)  # :644-645
        _req(G, False)
        with torch.no_grad():
            self.opt_g.step(grads)                                                   # :646
        return {"losses_discriminator_real": loss_d_real.item(),                     # :648-651
                "losses_discriminator_synthetic": loss_d_synth.item(),
                "losses_generator_synthetic": loss_g_adv.item(),
                "l1_losses_generator_synthetic": loss_l1.item(),
                "synthetic": synthetic.detach()}


# ---------------------------------------------------------------------------------------------- cycle step
class CycleTrainer:
    """Model.train_cycle inner loop (model.py:678-751) for CycleGAN / AttentionGAN with topography conditions."""

    def __init__(self, nets, model, add_identity_loss=False, py_seed=None):
        self.g_forward = GENERATOR_FORWARD[model]
        self.G_pp, self.G_pr = nets["pre_to_post_generator"], nets["post_to_pre_generator"]
        self.D_pre, self.D_post = nets["pre_discriminator"], nets["post_discriminator"]
        self.identity = add_identity_loss
        self.opt_g = Adam(_float_params(self.G_pp) + _float_params(self.G_pr))      # :112-114
        self.opt_d = Adam(_float_params(self.D_post) + _float_params(self.D_pre))   # :115-117
        self.pre_buffer, self.post_buffer = [], []
        self.rng = random.Random(py_seed) if py_seed is not None else random

    def _buffer(self, image, buf):
        """get_buffer_image, model.py:275-294"""
        image = image.detach()
        if len(buf) < 50:
            buf.append(image.clone())
            return image
        if self.rng.uniform(0, 1) > 0.5:
            idx = self.rng.randint(0, 49)
            old = buf[idx].clone()
            buf[idx] = image.clone()
            return old
        return image

    def step(self, input_stack, output_image):
        gf = self.g_forward
        for p in (self.G_pp, self.G_pr):
            _req(p, True)
        real_pre = input_stack                                                        # :680
        cond = input_stack[:, 3:].detach().clone()                                    # :683
        real_post = torch.cat((output_image, cond), 1)                                # :684
        synth_post = gf(self.G_pp, real_pre)                                          # :685
        synth_pre = gf(self.G_pr, real_post)                                          # :686
        synth_post = torch.cat((synth_post, cond), 1)                                 # :688
        synth_pre = torch.cat((synth_pre, cond), 1)                                   # :689
        rec_post = gf(self.G_pp, synth_pre)                                           # :690
        rec_pre = gf(self.G_pr, synth_post)                                           # :691
        _req(self.D_pre, False)
        _req(self.D_post, False)
        idt_post = idt_pre = 0
        if self.identity:                                                             # :702-704
            idt_post = F.l1_loss(gf(self.G_pp, real_post), real_post[:, :3]) * 5
            idt_pre = F.l1_loss(gf(self.G_pr, real_pre), real_pre[:, :3]) * 5
        pred = patchgan_forward(self.D_post, synth_post)
        loss_g_post = F.mse_loss(pred, torch.full(pred.shape, 1.0))                   # :706-707
        pred = patchgan_forward(self.D_pre, synth_pre)
        loss_g_pre = F.mse_loss(pred, torch.full(pred.shape, 1.0))                    # :708-709
        cyc_pre = F.l1_loss(rec_pre, real_pre[:, :3]) * 10                            # :710
        cyc_post = F.l1_loss(rec_post, real_post[:, :3]) * 10                         # :711
        loss_g = loss_g_post + loss_g_pre + cyc_pre + cyc_post + idt_post + idt_pre   # :712
        gparams = _float_params(self.G_pp) + _float_params(self.G_pr)
        grads = torch.autograd.grad(loss_g, gparams, allow_unused=True)
        for p in (self.G_pp, self.G_pr):
            _req(p, False)
        with torch.no_grad():
            self.opt_g.step(grads)                                                    # :714
        _req(self.D_pre, True)
        _req(self.D_post, True)
        sp = self._buffer(synth_pre, self.pre_buffer)                                 # :723
        spo = self._buffer(synth_post, self.post_buffer)                              # :724
        pr = patchgan_forward(self.D_pre, real_pre)
        l_real_pre = F.mse_loss(pr, torch.full(pr.shape, 1.0))                        # :726-727
        ps = patchgan_forward(self.D_pre, sp.detach())
        l_syn_pre = F.mse_loss(ps, torch.full(ps.shape, 0.0))                         # :728-729
        pr2 = patchgan_forward(self.D_post, real_post)
        l_real_post = F.mse_loss(pr2, torch.full(pr2.shape, 1.0))                     # :733-734
        ps2 = patchgan_forward(self.D_post, spo.detach())
        l_syn_post = F.mse_loss(ps2, torch.full(ps2.shape, 0.0))                      # :735-736
        dparams = _float_params(self.D_post) + _float_params(self.D_pre)
        total = (l_real_pre + l_syn_pre) * 0.5 + (l_real_post + l_syn_post) * 0.5     # :730-731, :737-738
        dgrads = torch.autograd.grad(total, dparams, allow_unused=True)
        _req(self.D_pre, False)
        _req(self.D_post, False)
        with torch.no_grad():
            self.opt_d.step(dgrads)                                                   # :739
        out = {"losses_generator_post": loss_g_post.item(), "losses_generator_pre": loss_g_pre.item(),
               "losses_pre_to_post_cycle": cyc_pre.item(), "losses_post_to_pre_cycle": cyc_post.item(),
               "losses_discriminator_pre_real": l_real_pre.item(), "losses_discriminator_post_real": l_real_post.item(),
               "losses_discriminator_pre_synthetic": l_syn_pre.item(),
               "losses_discriminator_post_synthetic": l_syn_post.item()}
        if self.identity:
            out["losses_identity_post"] = idt_post.item()
            out["losses_identity_pre"] = idt_pre.item()
        return out


# ---------------------------------------------------------------------------------------------- segmentation U-Net
def unet_plan(n_channels=3, n_classes=1):
    """UNet(bilinear=False): model_architectures.py:508-586 (construction order == registration order == apply order)"""
    p = []

    def double(prefix, cin, cout):
        p.extend([(prefix + ".0", (cout, cin, 3, 3), False, "conv"), (prefix + ".1", (cout,), False, "bn"),
                  (prefix + ".3", (cout, cout, 3, 3), False, "conv"), (prefix + ".4", (cout,), False, "bn")])

    double("inc.double_conv", n_channels, 64)
    for i, (cin, cout) in enumerate(((64, 128), (128, 256), (256, 512), (512, 1024)), 1):
        double(f"down{i}.maxpool_conv.1.double_conv", cin, cout)
    for i, (cin, cout) in enumerate(((1024, 512), (512, 256), (256, 128), (128, 64)), 1):
        p.append((f"up{i}.up", (cin, cin // 2, 2, 2), True, "convT"))
        double(f"up{i}.conv.double_conv", cin, cout)
    p.append(("outc.conv", (n_classes, 64, 1, 1), True, "conv"))
    return p


def init_unet(seed=None):
    """SegmentationModel.__init__ (segmentation_model.py:55): UNet().apply(initialise_weights)"""
    if seed is not None:
        torch.manual_seed(seed)
    return _init_from_plan(unet_plan())


def unet_forward(p, x):
    """UNet.forward in TRAINING mode -- the reference never calls .eval() on the segmentation model
    (segmentation_model.py:55-61, model.py:380-400), so BatchNorm uses batch statistics and updates its buffers."""
    def double(prefix, t):
        t = F.relu(_bn(p, prefix + ".1", F.conv2d(t, p[prefix + ".0.weight"], None, padding=1)))
        return F.relu(_bn(p, prefix + ".4", F.conv2d(t, p[prefix + ".3.weight"], None, padding=1)))

    xs = [double("inc.double_conv", x)]
    for i in range(1, 5):
        xs.append(double(f"down{i}.maxpool_conv.1.double_conv", F.max_pool2d(xs[-1], 2)))
    t = xs[4]
    for i in range(1, 5):
        up = F.conv_transpose2d(t, p[f"up{i}.up.weight"], p[f"up{i}.up.bias"], stride=2)
        skip = xs[4 - i]
        t = double(f"up{i}.conv.double_conv", torch.cat((skip, up), 1))  # sizes match for inputs divisible by 16
    return F.conv2d(t, p["outc.conv.weight"], p["outc.conv.bias"])


def segmentation_masks(p, generated, ground_truth):
    """model.py:397-400: both images rescaled from [-1, 1] to [0, 1], segmented, thresholded"""
    gt = torch.clamp((ground_truth + 1) * 0.5, min=0, max=1)
    gen = torch.clamp((generated + 1) * 0.5, min=0, max=1)
    return flood_mask(unet_forward(p, gen)), flood_mask(unet_forward(p, gt))


# ---------------------------------------------------------------------------------------------- integer work
def flood_mask(logits):
    """(sigmoid(x) > 0.5).float(): model.py:399-400, segmentation_model.py:244-248 (fp32, CPU)"""
    return (torch.sigmoid(logits.float()) > 0.5).float()


def confusion_counts(pred, truth):
    p, t = pred > 0.5, truth > 0.5
    return [int((p & t).sum()), int((p & ~t).sum()), int((~p & ~t).sum()), int((~p & t).sum())]


def synthetic_batch(step, batch, channels=9, size=256):
    """i.i.d. uniform [-1, 1], generator seeded with 1000 + step."""
    g = torch.Generator().manual_seed(1000 + step)
    x = torch.rand(batch, channels, size, size, generator=g) * 2 - 1
    y = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    return x, y
