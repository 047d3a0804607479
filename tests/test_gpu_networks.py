"""Network- and step-level parity of the CUDA engine against golden vectors recorded from the reference and
against the CPU oracle (fp32 mode: 1e-3 relative, the north-star tolerance)."""
import os
import random

import pytest
import torch

import stylegan_oracle as O
from test_oracle_golden import arch_from

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(arch, g_sd, d_sd):
    from models.GAN import Generator, Discriminator
    g = Generator(**arch.g_kwargs()).to(DEV)
    d = Discriminator(**arch.d_kwargs()).to(DEV)
    g.load_state_dict(g_sd)
    d.load_state_dict(d_sd)
    return g, d


def set_noise(g, noises):
    from models.CustomLayers import NoiseLayer
    layers = [m for m in g.g_synthesis.modules() if isinstance(m, NoiseLayer)]
    for m, n in zip(layers, noises):
        m.noise = n.to(DEV)


class FixedMixing:
    """Replays the reference's host RNG draws (z2, cut-off) recorded in the fixture."""

    def __init__(self, z2, cutoff, depth):
        self.z2, self.cutoff, self.depth = z2, cutoff, depth

    def __enter__(self):
        self._randn, self._random, self._randint = torch.randn, random.random, random.randint
        z2 = self.z2
        cutoff, cur = self.cutoff, 2 * (self.depth + 1)

        def fake_randn(*a, **k):
            torch.randn = self._randn
            return z2.clone()
        torch.randn = fake_randn
        random.random = lambda: 0.0 if cutoff != cur else 0.0
        random.randint = lambda a, b: cutoff
        return self

    def __exit__(self, *a):
        torch.randn, random.random, random.randint = self._randn, self._random, self._randint


@pytest.fixture(autouse=True)
def _fp32():
    import models
    from models import _ops
    _ops.clear_weight_cache()
    models.set_precision("fp32")
    yield
    models.set_precision("fp32")


@pytest.mark.parametrize("name", ["fb_tiny16_d0", "fb_tiny128_d5", "fb_tiny128_d4"])
def test_forward_backward_vs_reference_golden(golden_dir, name):
    from models import Losses
    fx = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    arch = arch_from(fx["arch"])
    depth, alpha = fx["depth"], fx["alpha"]
    g, d = build(arch, fx["g_init"], fx["d_init"])
    g.train(); d.train()
    set_noise(g, fx["noises"])
    with FixedMixing(fx["mix"][0], fx["mix"][1], depth):
        fake = g(fx["z"].to(DEV), depth, alpha)
    assert rel(fake, fx["fake"]) < 1e-3
    assert rel(g.truncation.avg_latent, fx["avg_after"]) < 1e-4
    assert rel(d(fake.detach(), depth, alpha), fx["logits_fake"]) < 1e-3
    real = fx["real"].to(DEV)
    assert rel(d(real, depth, alpha), fx["logits_real"]) < 1e-3
    loss = Losses.LogisticGAN(d)
    r1 = loss.R1Penalty(real, depth, alpha)
    assert abs(float(r1) - fx["r1"]) <= 2e-3 * abs(fx["r1"])
    d_loss = loss.dis_loss(real, fake.detach(), depth, alpha)
    assert abs(float(d_loss) - fx["d_loss"]) <= 1e-3 * abs(fx["d_loss"])
    names = [n for n, _ in d.named_parameters()]
    grads = torch.autograd.grad(d_loss, list(d.parameters()), allow_unused=True)
    for n, gr in zip(names, grads):
        if gr is None:
            assert n not in fx["d_grads"], n
            continue
        ref = fx["d_grads"][n]
        assert abs(float(gr.norm()) - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-8, n
        if ref["full"] is not None:
            assert rel(gr, ref["full"]) < 2e-3, n
    g_loss = loss.gen_loss(None, fake, depth, alpha)
    assert abs(float(g_loss) - fx["g_loss"]) <= 1e-3 * abs(fx["g_loss"])
    names = [n for n, _ in g.named_parameters()]
    grads = torch.autograd.grad(g_loss, list(g.parameters()), allow_unused=True)
    for n, gr in zip(names, grads):
        if gr is None:
            assert n not in fx["g_grads"], n
            continue
        ref = fx["g_grads"][n]
        assert abs(float(gr.norm()) - ref["norm"]) <= 3e-3 * ref["norm"] + 1e-8, n
        assert rel(gr.flatten()[:64], ref["head"]) < 5e-3, n


def _make_stylegan(arch, g_sd, d_sd):
    from models.GAN import StyleGAN
    gk, dk = arch.g_kwargs(), arch.d_kwargs()
    for k in ("resolution", "num_channels", "structure"):
        gk.pop(k), dk.pop(k)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)       # int beta_1 as in config.py:81
    sg = StyleGAN(structure=arch.structure, resolution=arch.resolution, num_channels=arch.num_channels,
                  latent_size=arch.latent_size, g_args=gk, d_args=dk, g_opt_args=opt, d_opt_args=dict(opt),
                  loss="logistic", use_ema=True, ema_decay=0.999, device=torch.device(DEV))
    sg.gen.load_state_dict(g_sd)
    sg.dis.load_state_dict(d_sd)
    sg.gen_shadow.load_state_dict(g_sd)
    from models import _ops
    _ops.clear_weight_cache()
    return sg


def test_training_trajectory_vs_reference_golden(golden_dir):
    """10 G+D steps (BASELINE config 1 in miniature): per-step losses match the reference run."""
    fx = torch.load(os.path.join(golden_dir, "traj_tiny16.pt"), weights_only=False)
    arch = arch_from(fx["arch"])
    sg = _make_stylegan(arch, fx["g_init"], fx["d_init"])
    sg.gen.train(); sg.dis.train(); sg.gen_shadow.train()
    set_noise(sg.gen, fx["noises"])
    for (depth, alpha), s in zip(fx["schedule"], fx["steps"]):
        z, img = s["z"].to(DEV), s["images"].to(DEV)
        with FixedMixing(s["mix_d"][0], s["mix_d"][1], depth):
            dl = sg.optimize_discriminator(z, img, depth, alpha)
        with FixedMixing(s["mix_g"][0], s["mix_g"][1], depth):
            gl = sg.optimize_generator(z, img, depth, alpha)
        assert abs(dl - s["d_loss"]) <= 2e-3 * abs(s["d_loss"]), (depth, alpha, dl, s["d_loss"])
        assert abs(gl - s["g_loss"]) <= 2e-3 * abs(s["g_loss"]), (depth, alpha, gl, s["g_loss"])
    # final state: see tests/test_oracle_golden.py for why this is loose (Adam beta1=0 turns round-off-level
    # gradients into +-lr steps; 10 steps * lr 0.003 on O(1) parameters)
    for k, v in fx["d_final"].items():
        assert rel(sg.dis.state_dict()[k], v) < 5e-2, k
    for k, v in fx["g_final"].items():
        assert rel(sg.gen.state_dict()[k], v) < 5e-2, k
    for k, v in fx["shadow_final"].items():
        if k != "truncation.avg_latent":
            assert rel(sg.gen_shadow.state_dict()[k], v) < 5e-2, k
    names = [n for n, _ in sg.gen.named_parameters()]
    for n, p in zip(names, sg.gen.parameters()):
        if n in fx["g_adam_step"]:
            assert int(sg.gen_optim.state[p]["step"]) == fx["g_adam_step"][n], n
            assert rel(sg.gen_optim.state[p]["exp_avg_sq"], fx["g_exp_avg_sq"][n]) < 5e-2, n
    # checkpoint layout round-trip (reference GAN.py:813-823 / train.py:102-126)
    sd = sg.gen_optim.state_dict()
    assert set(sd) == {"state", "param_groups"}
    sg.gen_optim.load_state_dict(sd)


def test_bf16_step_close_to_fp32_oracle():
    """bf16 engine vs fp32 oracle on one G+D step at real channel widths (parity unpinned: loose tolerance)."""
    import models
    torch.manual_seed(0)
    random.seed(0)
    arch = O.Arch(32, fmap_base=2048, fmap_max=128, latent_size=64, mapping_fmaps=64)
    gen = torch.Generator().manual_seed(0)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    for k in g_sd:
        if k.endswith("noise.weight"):
            g_sd[k] = torch.randn(g_sd[k].shape, generator=gen)
    B, depth, alpha = 8, 3, 0.5
    noises = [torch.randn(s, generator=gen) for s in O.noise_shapes(arch, B, depth)]
    z = torch.randn(B, arch.latent_size, generator=gen)
    z2 = torch.randn(B, arch.latent_size, generator=gen)
    real = torch.randn(B, 3, 32, 32, generator=gen)
    sd_g = {k: v.clone() for k, v in g_sd.items()}
    fake_ref = O.generator_forward(sd_g, z, depth, alpha, arch, z2=z2, mixing_cutoff=4, noises=noises)
    d_ref = O.d_loss_logistic(d_sd, real, fake_ref, depth, alpha, arch)
    for prec, tol in (("fp32", 1e-3), ("bf16", 8e-2)):
        models.set_precision(prec)
        from models import Losses, _ops
        _ops.clear_weight_cache()
        g, d = build(arch, g_sd, d_sd)
        g.train()
        set_noise(g, noises)
        with FixedMixing(z2, 4, depth):
            fake = g(z.to(DEV), depth, alpha)
        assert rel(fake, fake_ref) < tol, prec
        dl = Losses.LogisticGAN(d).dis_loss(real.to(DEV), fake_ref.to(DEV), depth, alpha)
        assert abs(float(dl) - float(d_ref)) < tol * abs(float(d_ref)) * 2, prec
