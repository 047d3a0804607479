"""Wider coverage of the drop-in networks on the GPU: every depth and fade-in alpha, 'fixed' structure, eval mode,
the non-default losses, conditional mode, checkpoint round trip (the pyramid the reference lacks)."""
import io
import os
import random

import pytest
import torch

import stylegan_oracle as O
from test_gpu_networks import FixedMixing, build, rel, set_noise

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32():
    import models
    from models import _ops
    _ops.clear_weight_cache()
    models.set_precision("fp32")
    yield
    models.set_precision("fp32")


def _init(arch, seed=0, batch=4):
    gen = torch.Generator().manual_seed(seed)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    for k in list(g_sd) + list(d_sd):
        sd = g_sd if k in g_sd else d_sd
        if k.endswith("noise.weight") or (k.endswith(".bias") and "init_block.bias" not in k):
            sd[k] = 0.3 * torch.randn(sd[k].shape, generator=gen)
    return g_sd, d_sd, gen


@pytest.mark.parametrize("depth,alpha", [(0, 1.0), (1, 0.0), (1, 0.5), (2, 1.0), (3, 0.25), (4, 0.5), (5, 0.75)])
def test_every_depth_and_alpha_vs_oracle(depth, alpha):
    arch = O.Arch(128, latent_size=32, mapping_fmaps=32, mapping_layers=2, fmap_base=512, fmap_max=32)
    g_sd, d_sd, gen = _init(arch)
    B = 4
    noises = [torch.randn(s, generator=gen) for s in O.noise_shapes(arch, B, depth)]
    z = torch.randn(B, arch.latent_size, generator=gen)
    z2 = torch.randn(B, arch.latent_size, generator=gen)
    cutoff = 1 + depth
    res = 2 ** (depth + 2)
    real = torch.randn(B, 3, res, res, generator=gen)
    sd_g = {k: v.clone() for k, v in g_sd.items()}
    fake_ref = O.generator_forward(sd_g, z, depth, alpha, arch, z2=z2, mixing_cutoff=cutoff, noises=noises)
    logit_ref = O.discriminator(d_sd, real, depth, alpha, arch)
    g, d = build(arch, g_sd, d_sd)
    g.train()
    set_noise(g, noises)
    with FixedMixing(z2, cutoff, depth):
        fake = g(z.to(DEV), depth, alpha)
    assert fake.shape == fake_ref.shape and fake.dtype == torch.float32
    assert rel(fake, fake_ref) < 1e-3
    assert rel(d(real.to(DEV), depth, alpha), logit_ref) < 1e-3
    # progressive down-sampling of the real batch (reference GAN.py:557-589)
    from models.GAN import StyleGAN
    full = torch.randn(B, 3, 128, 128, generator=gen)
    sg = StyleGAN.__new__(StyleGAN)
    sg.structure, sg.depth = "linear", arch.depth
    got = sg._StyleGAN__progressive_down_sampling(full.to(DEV), depth, alpha)
    assert rel(got, O.progressive_down_sampling(full, arch, depth, alpha)) < 1e-5


def test_fixed_structure_and_eval_mode():
    arch = O.Arch(32, latent_size=32, mapping_fmaps=32, mapping_layers=2, fmap_base=256, fmap_max=32,
                  structure="fixed")
    g_sd, d_sd, gen = _init(arch, seed=3)
    B = 4
    depth = arch.depth - 1
    noises = [torch.randn(s, generator=gen) for s in O.noise_shapes(arch, B, depth)]
    z = torch.randn(B, arch.latent_size, generator=gen)
    g, d = build(arch, g_sd, d_sd)
    g.eval()                       # eval: no style mixing, no truncation, no avg update (reference GAN.py:274)
    set_noise(g, noises)
    avg0 = g.truncation.avg_latent.clone()
    fake = g(z.to(DEV), depth, 1.0)
    fake_ref = O.generator_forward({k: v.clone() for k, v in g_sd.items()}, z, depth, 1.0, arch, training=False,
                                   noises=noises)
    assert rel(fake, fake_ref) < 1e-3
    assert torch.equal(g.truncation.avg_latent, avg0)
    assert rel(d(fake_ref.to(DEV), depth, 1.0), O.discriminator(d_sd, fake_ref, depth, 1.0, arch)) < 1e-3
    with pytest.raises(AssertionError):
        g.g_synthesis(torch.zeros(B, arch.num_layers, 512, device=DEV), depth=arch.depth)


@pytest.mark.parametrize("res,depth,alpha,fmap_base,fmap_max", [(16, 2, 0.5, 128, 32), (16, 0, 1.0, 128, 32),
                                                                 (128, 5, 0.3, 512, 32)])
def test_shared_real_pass_equals_two_pass(res, depth, alpha, fmap_base, fmap_max):
    """LogisticGAN.dis_loss with one shared D(real) graph for the logistic term and the R1 penalty (the default) against
    the reference's two-pass formulation (Losses.py:203,214): same loss, same parameter gradients."""
    import models
    from models.GAN import Discriminator
    from models.Losses import LogisticGAN
    torch.manual_seed(res + depth)
    d = Discriminator(resolution=res, fmap_base=fmap_base, fmap_max=fmap_max, blur_filter=[1, 2, 1]).to(DEV)
    with torch.no_grad():
        for n, p in d.named_parameters():
            if n.endswith(".bias"):
                p.copy_(0.1 * torch.randn_like(p))
    d.train()
    L = LogisticGAN(d)
    r = 2 ** (depth + 2)
    real, fake = torch.randn(4, 3, r, r, device=DEV), torch.randn(4, 3, r, r, device=DEV)
    params = list(d.parameters())
    out = []
    try:
        for shared in (False, True):
            models.set_shared_r1(shared)
            loss = L.dis_loss(real, fake, depth, alpha)
            out.append((float(loss.detach()), torch.autograd.grad(loss, params, allow_unused=True)))
    finally:
        models.set_shared_r1(True)
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[0][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert (a is None) == (b is None)
        if a is not None:
            assert rel(b, a) < 1e-4


@pytest.mark.parametrize("loss_name", ["hinge", "relativistic-hinge"])
def test_other_losses_vs_oracle(loss_name):
    from models import Losses
    arch = O.Arch(16, latent_size=32, mapping_fmaps=32, mapping_layers=2, fmap_base=128, fmap_max=32)
    g_sd, d_sd, gen = _init(arch, seed=5)
    B, depth, alpha = 4, 2, 0.5
    real = torch.randn(B, 3, 16, 16, generator=gen)
    fake = torch.randn(B, 3, 16, 16, generator=gen)
    g, d = build(arch, g_sd, d_sd)
    cls = {"hinge": Losses.HingeGAN, "relativistic-hinge": Losses.RelativisticAverageHingeGAN}[loss_name]
    L = cls(d)
    dl = L.dis_loss(real.to(DEV), fake.to(DEV), depth, alpha)
    gl = L.gen_loss(real.to(DEV), fake.to(DEV), depth, alpha)
    if loss_name == "hinge":
        dl_o, gl_o = O.d_loss_hinge(d_sd, real, fake, depth, alpha, arch), O.g_loss_hinge(d_sd, fake, depth, alpha, arch)
    else:
        dl_o = O.d_loss_rahinge(d_sd, real, fake, depth, alpha, arch)
        gl_o = O.g_loss_rahinge(d_sd, real, fake, depth, alpha, arch)
    assert abs(float(dl) - float(dl_o)) < 1e-3 * abs(float(dl_o)) + 1e-5
    assert abs(float(gl) - float(gl_o)) < 1e-3 * abs(float(gl_o)) + 1e-5
    grads = torch.autograd.grad(dl, [p for p in d.parameters()], allow_unused=True)
    assert any(gr is not None and float(gr.abs().sum()) > 0 for gr in grads)


def test_conditional_mode_runs_and_checkpoint_roundtrip(tmp_path):
    """Conditional G/D (class embedding concatenation, reference GAN.py:233-236,326-369) and the five checkpoint files."""
    from models.GAN import StyleGAN
    torch.manual_seed(0)
    random.seed(0)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)
    g_args = dict(latent_size=32, mapping_fmaps=32, mapping_layers=2, fmap_base=128, fmap_max=32, blur_filter=[1, 2, 1])
    d_args = dict(fmap_base=128, fmap_max=32, blur_filter=[1, 2, 1])
    sg = StyleGAN("linear", 16, 3, 32, g_args, d_args, opt, dict(opt), conditional=True, n_classes=5,
                  loss="conditional-loss", use_ema=True, device=torch.device(DEV))
    B, depth, alpha = 4, 1, 0.5
    labels = torch.tensor([0, 1, 2, 4], device=DEV)
    img = torch.randn(B, 3, 16, 16, device=DEV)
    z = torch.randn(B, 32, device=DEV)
    dl = sg.optimize_discriminator(z, img, depth, alpha, labels)
    gl = sg.optimize_generator(z, img, depth, alpha, labels)
    assert dl == dl and gl == gl                      # finite
    # checkpoint layout (reference GAN.py:806-823): plain state_dicts that load back strictly
    files = {}
    for name, obj in (("GEN", sg.gen), ("DIS", sg.dis), ("GEN_SHADOW", sg.gen_shadow), ("GEN_OPTIM", sg.gen_optim),
                      ("DIS_OPTIM", sg.dis_optim)):
        f = tmp_path / ("GAN_%s_1_1.pth" % name)
        torch.save(obj.state_dict(), f)
        files[name] = f
    sg2 = StyleGAN("linear", 16, 3, 32, g_args, d_args, opt, dict(opt), conditional=True, n_classes=5,
                   loss="conditional-loss", use_ema=True, device=torch.device(DEV))
    sg2.gen.load_state_dict(torch.load(files["GEN"]))
    sg2.dis.load_state_dict(torch.load(files["DIS"]))
    sg2.gen_shadow.load_state_dict(torch.load(files["GEN_SHADOW"]))
    sg2.gen_optim.load_state_dict(torch.load(files["GEN_OPTIM"]))
    sg2.dis_optim.load_state_dict(torch.load(files["DIS_OPTIM"]))
    for a, b in zip(sg.gen.parameters(), sg2.gen.parameters()):
        assert torch.equal(a, b)
    p0 = next(iter(sg2.dis_optim.state))
    assert float(sg2.dis_optim.state[p0]["step"]) >= 0
    # and training continues from the restored state
    sg2.optimize_discriminator(z, img, depth, alpha, labels)


def test_bf16_training_is_stable_and_learns():
    """30 bf16 G+D steps at 32x32 with real channel widths for the tensor-core kernels: losses stay finite, the
    discriminator loss moves, parameters change, EMA shadow follows (a smoke test of the perf path as a whole)."""
    import models
    from models.GAN import StyleGAN
    from models import CustomLayers
    models.set_precision("bf16")
    torch.manual_seed(0)
    random.seed(0)
    CustomLayers.manual_seed(0)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)
    g_args = dict(latent_size=64, mapping_fmaps=64, mapping_layers=2, fmap_base=1024, fmap_max=128,
                  blur_filter=[1, 2, 1])
    d_args = dict(fmap_base=1024, fmap_max=128, blur_filter=[1, 2, 1])
    sg = StyleGAN("linear", 32, 3, 64, g_args, d_args, opt, dict(opt), loss="logistic", use_ema=True,
                  device=torch.device(DEV))
    w0 = sg.gen.g_synthesis.blocks[0].conv1.weight.detach().clone()
    s0 = sg.gen_shadow.g_synthesis.blocks[0].conv1.weight.detach().clone()
    B, depth = 8, 3
    d_losses, g_losses = [], []
    base = torch.randn(1, 3, 32, 32, device=DEV)
    for i in range(30):
        img = (base + 0.1 * torch.randn(B, 3, 32, 32, device=DEV)).clamp(-1, 1)
        z = torch.randn(B, 64, device=DEV)
        d_losses.append(sg.optimize_discriminator(z, img, depth, min(1.0, (i + 1) / 10)))
        g_losses.append(sg.optimize_generator(z, img, depth, min(1.0, (i + 1) / 10)))
    # GAN trajectories are chaotic (and the reductions use atomics), so only robust facts are asserted
    assert all(l == l and abs(l) < 1e4 for l in d_losses + g_losses), (d_losses, g_losses)
    assert max(d_losses) - min(d_losses) > 1e-3
    w1 = sg.gen.g_synthesis.blocks[0].conv1.weight.detach()
    assert float((w1 - w0).abs().max()) > 1e-3
    s1 = sg.gen_shadow.g_synthesis.blocks[0].conv1.weight.detach()
    assert 0 < float((s1 - s0).abs().max()) < float((w1 - w0).abs().max())
    out = sg.gen_shadow(torch.randn(4, 64, device=DEV), depth, 1.0)
    assert out.shape == (4, 3, 32, 32) and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-3), ("bf16", 8e-2)])
def test_full_width_step_vs_oracle(prec, tol):
    """One G+D step at the reference's REAL channel widths (512/512/512/512/256 at 4..64 px, batch 4, depth 4,
    alpha 0.5, logistic + R1, EMA) against the CPU oracle: losses and a few updated parameters."""
    import models
    from models.GAN import StyleGAN
    from models.CustomLayers import NoiseLayer
    from models import _ops
    arch = O.Arch(64)
    gen = torch.Generator().manual_seed(11)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    for k in g_sd:
        if k.endswith("noise.weight"):
            g_sd[k] = 0.5 * torch.randn(g_sd[k].shape, generator=gen)
    B, depth, alpha = 4, 4, 0.5
    noises = [torch.randn(s, generator=gen) for s in O.noise_shapes(arch, B, depth)]
    z = torch.randn(B, 512, generator=gen)
    img = torch.randn(B, 3, 64, 64, generator=gen)
    torch.manual_seed(5); random.seed(5)
    st = O.TrainState(arch, g_sd, d_sd)
    mix_d = O.draw_mixing(arch, z.shape, depth)
    mix_g = O.draw_mixing(arch, z.shape, depth)
    dl_o, gl_o = O.train_step(st, img, z, depth, alpha, mix_d, mix_g, noises, noises)

    models.set_precision(prec)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)
    sg = StyleGAN("linear", 64, 3, 512, dict(latent_size=512, mapping_layers=4, blur_filter=[1, 2, 1],
                                             truncation_psi=0.7, truncation_cutoff=8),
                  dict(use_wscale=True, blur_filter=[1, 2, 1]), opt, dict(opt), loss="logistic", use_ema=True,
                  device=torch.device(DEV))
    sg.gen.load_state_dict(g_sd); sg.dis.load_state_dict(d_sd); sg.gen_shadow.load_state_dict(g_sd)
    _ops.clear_weight_cache()
    for m, n in zip([m for m in sg.gen.g_synthesis.modules() if isinstance(m, NoiseLayer)], noises):
        m.noise = n.to(DEV)
    torch.manual_seed(5); random.seed(5)
    dl = sg.optimize_discriminator(z.to(DEV), img.to(DEV), depth, alpha)
    gl = sg.optimize_generator(z.to(DEV), img.to(DEV), depth, alpha)
    assert abs(dl - dl_o) <= tol * abs(dl_o), (prec, dl, dl_o)
    assert abs(gl - gl_o) <= tol * abs(gl_o) + (0.02 if prec == "bf16" else 0.0), (prec, gl, gl_o)
    if prec == "fp32":
        # after one Adam step with beta1 = 0 every touched parameter moved by ~lr*sign(g): compare the directions
        for k in ("final_block.dense1.weight", "blocks.2.conv0.bias", "from_rgb.1.weight"):
            mine = sg.dis.state_dict()[k].cpu() - d_sd[k]
            ref = st.d[k] - d_sd[k]
            agree = float(((mine * ref) > 0).float().mean())
            assert agree > 0.9, (k, agree)      # round-off-sized gradients have random sign


def test_standard_and_conditional_losses_vs_reference_golden(golden_dir):
    """StandardGAN.dis_loss and the CONDITIONAL generator / discriminator with ConditionalGANLoss against
    values recorded from the live reference (oracle/make_golden.py --losses; reference models/Losses.py:54-133,
    GAN.py:233-236, 326-369, 395-437): forward image, logits, both losses and every parameter-gradient norm."""
    from models import Losses
    from models.GAN import Generator, Discriminator
    from test_gpu_networks import FixedMixing, set_noise
    from test_oracle_golden import arch_from
    fx = torch.load(os.path.join(golden_dir, "losses2_tiny16_d2.pt"), weights_only=False)
    arch = arch_from(fx["arch"])
    depth, alpha, K = fx["depth"], fx["alpha"], fx["n_classes"]
    real = fx["real"].to(DEV)
    # --- StandardGAN on the unconditional discriminator
    d = Discriminator(**arch.d_kwargs()).to(DEV)
    d.load_state_dict(fx["d_init"])
    d.train()
    dl = Losses.StandardGAN(d).dis_loss(real, fx["fake_in"].to(DEV), depth, alpha)
    ref = fx["standard"]
    assert abs(float(dl) - ref["d_loss"]) <= 1e-3 * abs(ref["d_loss"])
    grads = torch.autograd.grad(dl, list(d.parameters()), allow_unused=True)
    for (n, _), g in zip(d.named_parameters(), grads):
        if g is None:
            assert n not in ref["d_grad_norms"], n
        else:
            assert abs(float(g.norm()) - ref["d_grad_norms"][n]) <= 3e-3 * ref["d_grad_norms"][n] + 1e-9, n
    # gen_loss: the reference's own is not executable (Losses.py:131); ours is the evident intent, BCE against ones
    gl = Losses.StandardGAN(d).gen_loss(None, fx["fake_in"].to(DEV), depth, alpha)
    want = torch.nn.functional.binary_cross_entropy_with_logits(d(fx["fake_in"].to(DEV), depth, alpha).squeeze(),
                                                                torch.ones(fx["batch"], device=DEV))
    assert abs(float(gl) - float(want)) < 1e-6
    # --- conditional networks
    c = fx["conditional"]
    g = Generator(conditional=True, n_classes=K, **arch.g_kwargs()).to(DEV)
    dc = Discriminator(conditional=True, n_classes=K, **arch.d_kwargs()).to(DEV)
    assert list(g.state_dict().keys()) == list(c["g_init"].keys())
    assert list(dc.state_dict().keys()) == list(c["d_init"].keys())
    g.load_state_dict(c["g_init"]); dc.load_state_dict(c["d_init"])
    g.train(); dc.train()
    set_noise(g, c["noises"])
    labels = c["labels"].to(DEV)
    with FixedMixing(c["mix"][0], c["mix"][1], depth):
        fake = g(c["z"].to(DEV), depth, alpha, labels)
    assert rel(fake, c["fake"]) < 1e-3
    assert rel(dc(real, depth, alpha, labels_in=labels), c["logits_real"]) < 1e-3
    CL = Losses.ConditionalGANLoss(dc)
    cdl = CL.dis_loss(real, c["fake"].to(DEV), labels, depth, alpha)
    assert abs(float(cdl) - c["d_loss"]) <= 1e-3 * abs(c["d_loss"])
    grads = torch.autograd.grad(cdl, list(dc.parameters()), allow_unused=True)
    for (n, _), gr in zip(dc.named_parameters(), grads):
        if gr is None:
            assert n not in c["d_grad_norms"], n
        else:
            assert abs(float(gr.norm()) - c["d_grad_norms"][n]) <= 3e-3 * c["d_grad_norms"][n] + 1e-9, n
    cgl = CL.gen_loss(None, fake, labels, depth, alpha)
    assert abs(float(cgl) - c["g_loss"]) <= 1e-3 * abs(c["g_loss"])
    grads = torch.autograd.grad(cgl, list(g.parameters()), allow_unused=True)
    for (n, _), gr in zip(g.named_parameters(), grads):
        if gr is None:
            assert n not in c["g_grad_norms"], n
        else:
            assert abs(float(gr.norm()) - c["g_grad_norms"][n]) <= 5e-3 * c["g_grad_norms"][n] + 1e-9, n
