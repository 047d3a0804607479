"""Generate golden vectors by EXECUTING the unmodified reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Writes small ``tests/golden/*.pt`` fixtures (committed).  Each fixture holds the initial ``state_dict``s, every
host-side random draw the reference made (latents, second latents + mixing cut-offs, images, per-layer noise) and
the reference's outputs, so that both the oracle and the CUDA engine can be replayed against it without the
reference being importable.  Test infrastructure; never imported by the product path.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SG_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, HERE)

from models.GAN import StyleGAN  # noqa: E402  (the reference)
from models.CustomLayers import NoiseLayer  # noqa: E402
import stylegan_oracle as O  # noqa: E402


def seed_all(s):
    torch.manual_seed(s)
    random.seed(s)
    np.random.seed(s)


def build_reference(arch: O.Arch, loss="logistic", use_ema=True):
    gk = arch.g_kwargs()
    dk = arch.d_kwargs()
    for k in ("resolution", "num_channels", "structure"):
        gk.pop(k, None)
        dk.pop(k, None)
    latent = gk.pop("latent_size")
    gk["latent_size"] = latent
    opt = dict(learning_rate=0.003, beta_1=0.0, beta_2=0.99, eps=1e-8)   # config.py:79-92 (beta_1 as float)
    return StyleGAN(structure=arch.structure, resolution=arch.resolution, num_channels=arch.num_channels,
                    latent_size=latent, g_args=gk, d_args=dk, g_opt_args=opt, d_opt_args=dict(opt),
                    loss=loss, use_ema=use_ema, ema_decay=0.999, device=torch.device("cpu"))


def randomize_zero_inits(sg):
    """Noise weights and all biases start at zero in the reference; perturb them so those paths are exercised."""
    with torch.no_grad():
        for net in (sg.gen, sg.dis):
            for n, p in net.named_parameters():
                if n.endswith("noise.weight"):
                    p.copy_(torch.randn_like(p))
                elif n.endswith(".bias") and "init_block.bias" not in n:
                    p.copy_(0.1 * torch.randn_like(p))
        if sg.use_ema:
            sg.gen_shadow.load_state_dict(sg.gen.state_dict())


def set_noise(gen, batch):
    """Fix every NoiseLayer's noise (CustomLayers.py:194-198).  Returns the list in module order."""
    noises = []
    res = 4
    layers = [m for m in gen.g_synthesis.modules() if isinstance(m, NoiseLayer)]
    for i, m in enumerate(layers):
        r = 4 * 2 ** (i // 2)
        n = torch.randn(batch, 1, r, r)
        m.noise = n
        noises.append(n)
    return noises


def clone_sd(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def record_mixing(arch, z_shape, depth):
    """Peek the host RNG draws the next training-mode Generator.forward will make, then rewind."""
    ts, ps = torch.get_rng_state(), random.getstate()
    z2, cutoff = O.draw_mixing(arch, z_shape, depth)
    torch.set_rng_state(ts)
    random.setstate(ps)
    return z2, cutoff


def golden_trajectory(name, arch, batch, schedule, seed=0, outdir=None):
    """`schedule` = list of (depth, alpha) G+D steps."""
    seed_all(seed)
    sg = build_reference(arch)
    randomize_zero_inits(sg)
    sg.gen.train(); sg.dis.train(); sg.gen_shadow.train()
    noises = set_noise(sg.gen, batch)
    fx = dict(arch=arch.__dict__.copy(), batch=batch, schedule=schedule,
              g_init=clone_sd(sg.gen.state_dict()), d_init=clone_sd(sg.dis.state_dict()),
              noises=noises, steps=[])
    for depth, alpha in schedule:
        images = torch.randn(batch, arch.num_channels, arch.resolution, arch.resolution)
        z = torch.randn(batch, arch.latent_size)
        mix_d = record_mixing(arch, z.shape, depth)
        d_loss = sg.optimize_discriminator(z, images, depth, alpha)
        mix_g = record_mixing(arch, z.shape, depth)
        g_loss = sg.optimize_generator(z, images, depth, alpha)
        fx["steps"].append(dict(images=images, z=z, mix_d=mix_d, mix_g=mix_g, d_loss=d_loss, g_loss=g_loss))
    fx["g_final"] = clone_sd(sg.gen.state_dict())
    fx["d_final"] = clone_sd(sg.dis.state_dict())
    fx["shadow_final"] = clone_sd(sg.gen_shadow.state_dict())
    g_names = [n for n, _ in sg.gen.named_parameters()]
    d_names = [n for n, _ in sg.dis.named_parameters()]
    gs, ds = sg.gen_optim.state_dict()["state"], sg.dis_optim.state_dict()["state"]
    fx["g_exp_avg_sq"] = {g_names[i]: s["exp_avg_sq"].clone() for i, s in gs.items()}
    fx["d_exp_avg_sq"] = {d_names[i]: s["exp_avg_sq"].clone() for i, s in ds.items()}
    fx["g_adam_step"] = {g_names[i]: int(s["step"]) for i, s in gs.items()}
    fx["d_adam_step"] = {d_names[i]: int(s["step"]) for i, s in ds.items()}
    out = os.path.join(outdir or os.path.join(ROOT, "tests", "golden"), name + ".pt")
    torch.save(fx, out)
    print(name, "->", out, "%.1f KB" % (os.path.getsize(out) / 1024),
          "losses", [(round(s["d_loss"], 4), round(s["g_loss"], 4)) for s in fx["steps"]][:3], "...")


def golden_forward_backward(name, arch, batch, depth, alpha, seed=1):
    """Single-point forward/backward/R1 fixture (gradients kept as per-tensor norms + a few full tensors)."""
    seed_all(seed)
    sg = build_reference(arch)
    randomize_zero_inits(sg)
    sg.gen.train(); sg.dis.train()
    noises = set_noise(sg.gen, batch)
    z = torch.randn(batch, arch.latent_size)
    res = 2 ** (depth + 2)
    real = torch.randn(batch, arch.num_channels, res, res)
    mix = record_mixing(arch, z.shape, depth)
    g_init = clone_sd(sg.gen.state_dict())
    d_init = clone_sd(sg.dis.state_dict())
    fake = sg.gen(z, depth, alpha)
    avg_after = sg.gen.truncation.avg_latent.clone() if sg.gen.truncation is not None else None
    logits_fake = sg.dis(fake, depth, alpha)
    logits_real = sg.dis(real, depth, alpha)
    r1 = sg.loss.R1Penalty(real, depth, alpha)
    d_loss = sg.loss.dis_loss(real, fake.detach(), depth, alpha)
    d_params = [p for p in sg.dis.parameters()]
    d_grads = torch.autograd.grad(d_loss, d_params, allow_unused=True)
    g_loss = sg.loss.gen_loss(None, fake, depth, alpha)
    g_params = [p for p in sg.gen.parameters()]
    g_grads = torch.autograd.grad(g_loss, g_params, allow_unused=True)
    gn = [n for n, _ in sg.gen.named_parameters()]
    dn = [n for n, _ in sg.dis.named_parameters()]

    def pack(names, grads):
        out = {}
        for n, g in zip(names, grads):
            if g is None:
                continue
            out[n] = dict(norm=float(g.norm()), full=g.clone() if g.numel() <= 4096 else None,
                          head=g.flatten()[:64].clone())
        return out
    fx = dict(arch=arch.__dict__.copy(), batch=batch, depth=depth, alpha=alpha, g_init=g_init, d_init=d_init,
              noises=noises[:2 * (depth + 1)], z=z, real=real, mix=mix, fake=fake.detach(), avg_after=avg_after,
              logits_fake=logits_fake.detach(), logits_real=logits_real.detach(), r1=float(r1),
              d_loss=float(d_loss), g_loss=float(g_loss), d_grads=pack(dn, d_grads), g_grads=pack(gn, g_grads))
    out = os.path.join(ROOT, "tests", "golden", name + ".pt")
    torch.save(fx, out)
    print(name, "->", out, "%.1f KB" % (os.path.getsize(out) / 1024), "d_loss %.5f g_loss %.5f r1 %.5f" %
          (fx["d_loss"], fx["g_loss"], fx["r1"]))


def golden_other_losses(name, arch, batch, depth, alpha, seed=3):
    """Hinge and relativistic-average hinge (models/Losses.py:136-189) evaluated by the reference's own loss classes
    on its Discriminator: losses plus the gradient norms wrt every discriminator parameter."""
    from models.Losses import HingeGAN, RelativisticAverageHingeGAN
    seed_all(seed)
    sg = build_reference(arch)
    randomize_zero_inits(sg)
    sg.dis.train()
    res = 2 ** (depth + 2)
    real = torch.randn(batch, arch.num_channels, res, res)
    fake = torch.randn(batch, arch.num_channels, res, res)
    dn = [n for n, _ in sg.dis.named_parameters()]
    fx = dict(arch=arch.__dict__.copy(), batch=batch, depth=depth, alpha=alpha, d_init=clone_sd(sg.dis.state_dict()),
              real=real, fake=fake, losses={})
    for key, cls in (("hinge", HingeGAN), ("relativistic-hinge", RelativisticAverageHingeGAN)):
        L = cls(sg.dis)
        dl = L.dis_loss(real, fake, depth, alpha)
        gl = L.gen_loss(real, fake, depth, alpha)
        dg = torch.autograd.grad(dl, list(sg.dis.parameters()), allow_unused=True)
        gg = torch.autograd.grad(gl, list(sg.dis.parameters()), allow_unused=True)
        fx["losses"][key] = dict(d_loss=float(dl.detach()), g_loss=float(gl.detach()),
                                 d_grad_norms={n: float(g.norm()) for n, g in zip(dn, dg) if g is not None},
                                 g_grad_norms={n: float(g.norm()) for n, g in zip(dn, gg) if g is not None})
    out = os.path.join(ROOT, "tests", "golden", name + ".pt")
    torch.save(fx, out)
    print(name, "->", out, "%.1f KB" % (os.path.getsize(out) / 1024),
          {k: (round(v["d_loss"], 5), round(v["g_loss"], 5)) for k, v in fx["losses"].items()})


class _FixedMixing:
    """Make the reference's next training-mode Generator.forward draw (z2, cutoff) (GAN.py:282-290)."""

    def __init__(self, z2, cutoff):
        self.z2, self.cutoff = z2, cutoff

    def __enter__(self):
        self._randn, self._random, self._randint = torch.randn, random.random, random.randint
        z2, me = self.z2, self

        def fake_randn(*a, **k):
            torch.randn = me._randn
            return z2.clone()
        torch.randn = fake_randn
        random.random = lambda: 0.0
        random.randint = lambda a, b: self.cutoff

    def __exit__(self, *a):
        torch.randn, random.random, random.randint = self._randn, self._random, self._randint


def golden_full_width(name, res, batch, seed, alpha=0.5):
    """BASELINE configs 2/3 at the reference's REAL channel widths, top depth, small batch: the inputs are
    O.seeded_case(arch, batch, depth, seed) (re-derivable, not stored); stored are the reference's outputs: the image
    at the top depth (alpha) and at every lower depth (alpha 1), both logit vectors, R1, both losses, and for every
    parameter gradient its norm and a 64-value head (full tensor when <= 4096 elements)."""
    arch = O.Arch(res)
    depth = arch.depth - 1
    case = O.seeded_case(arch, batch, depth, seed)
    seed_all(seed)
    sg = build_reference(arch)
    sg.gen.load_state_dict(case["g_sd"]); sg.dis.load_state_dict(case["d_sd"])
    sg.gen.train(); sg.dis.train()
    layers = [m for m in sg.gen.g_synthesis.modules() if isinstance(m, NoiseLayer)]
    for m, n in zip(layers, case["noises"]):
        m.noise = n
    z, real = case["z"], case["real"]
    lower = {}
    with torch.no_grad():
        for d_ in range(depth):
            # the truncation average must not drift between these probes: restore it each time
            avg0 = sg.gen.truncation.avg_latent.clone()
            with _FixedMixing(case["z2"], min(case["cutoff"], 2 * (d_ + 1))):
                lower[d_] = sg.gen(z, d_, 1.0).clone()
            sg.gen.truncation.avg_latent.copy_(avg0)
    with _FixedMixing(case["z2"], case["cutoff"]):
        fake = sg.gen(z, depth, alpha)
    avg_after = sg.gen.truncation.avg_latent.clone()
    logits_fake = sg.dis(fake, depth, alpha)
    logits_real = sg.dis(real, depth, alpha)
    r1 = sg.loss.R1Penalty(real, depth, alpha)
    d_loss = sg.loss.dis_loss(real, fake.detach(), depth, alpha)
    d_grads = torch.autograd.grad(d_loss, list(sg.dis.parameters()), allow_unused=True)
    g_loss = sg.loss.gen_loss(None, fake, depth, alpha)
    g_grads = torch.autograd.grad(g_loss, list(sg.gen.parameters()), allow_unused=True)
    gn = [n for n, _ in sg.gen.named_parameters()]
    dn = [n for n, _ in sg.dis.named_parameters()]

    def pack(names, grads):
        out = {}
        for n, g in zip(names, grads):
            if g is None:
                continue
            out[n] = dict(norm=float(g.norm()), full=g.clone() if g.numel() <= 4096 else None,
                          head=g.flatten()[:64].clone())
        return out
    fx = dict(arch=arch.__dict__.copy(), res=res, batch=batch, depth=depth, alpha=alpha, seed=seed,
              fake=fake.detach().clone(), lower=lower, avg_after=avg_after, logits_fake=logits_fake.detach().clone(),
              logits_real=logits_real.detach().clone(), r1=float(r1), d_loss=float(d_loss), g_loss=float(g_loss),
              d_grads=pack(dn, d_grads), g_grads=pack(gn, g_grads))
    out = os.path.join(ROOT, "tests", "golden", name + ".pt")
    torch.save(fx, out)
    print(name, "->", out, "%.1f KB" % (os.path.getsize(out) / 1024), "d_loss %.5f g_loss %.5f r1 %.5f" %
          (fx["d_loss"], fx["g_loss"], fx["r1"]))


def golden_standard_and_conditional(name, arch, batch, depth, alpha, n_classes=4, seed=5):
    """The two remaining loss classes of the reference evaluated by the reference itself (models/Losses.py:54-133):
    StandardGAN.dis_loss on the unconditional discriminator (its gen_loss unpacks a 3-tuple the discriminator never
    returns, Losses.py:131, and cannot be executed), and ConditionalGANLoss.dis_loss / gen_loss on the CONDITIONAL
    generator + discriminator (class embeddings, GAN.py:233-236, 326-369, 395-437), with per-parameter gradient norms."""
    from models.Losses import StandardGAN, ConditionalGANLoss
    from models.GAN import Generator, Discriminator
    seed_all(seed)
    sg = build_reference(arch)
    randomize_zero_inits(sg)
    sg.dis.train()
    res = 2 ** (depth + 2)
    real = torch.randn(batch, arch.num_channels, res, res)
    fake_in = torch.randn(batch, arch.num_channels, res, res)
    dn = [n for n, _ in sg.dis.named_parameters()]
    L = StandardGAN(sg.dis)
    dl = L.dis_loss(real, fake_in, depth, alpha)
    dg = torch.autograd.grad(dl, list(sg.dis.parameters()), allow_unused=True)
    fx = dict(arch=arch.__dict__.copy(), batch=batch, depth=depth, alpha=alpha, n_classes=n_classes,
              d_init=clone_sd(sg.dis.state_dict()), real=real, fake_in=fake_in,
              standard=dict(d_loss=float(dl.detach()),
                            d_grad_norms={n: float(g.norm()) for n, g in zip(dn, dg) if g is not None}))
    # conditional networks
    gk, dk = arch.g_kwargs(), arch.d_kwargs()
    gen = Generator(conditional=True, n_classes=n_classes, **gk)
    dis = Discriminator(conditional=True, n_classes=n_classes, **dk)
    with torch.no_grad():
        for net in (gen, dis):
            for n, p in net.named_parameters():
                if n.endswith("noise.weight"):
                    p.copy_(torch.randn_like(p))
                elif n.endswith(".bias") and "init_block.bias" not in n:
                    p.copy_(0.1 * torch.randn_like(p))
    gen.train(); dis.train()
    noises = set_noise(gen, batch)
    z = torch.randn(batch, arch.latent_size)
    labels = torch.randint(0, n_classes, (batch,))
    mix = record_mixing(arch, (batch, 2 * arch.latent_size), depth)       # the conditional latent is [z | embedding]
    g_init, dc_init = clone_sd(gen.state_dict()), clone_sd(dis.state_dict())
    fake = gen(z, depth, alpha, labels)
    CL = ConditionalGANLoss(dis)
    logits_real = dis(real, depth, alpha, labels_in=labels)
    cdl = CL.dis_loss(real, fake.detach(), labels, depth, alpha)
    cdg = torch.autograd.grad(cdl, list(dis.parameters()), allow_unused=True)
    cgl = CL.gen_loss(None, fake, labels, depth, alpha)
    cgg = torch.autograd.grad(cgl, list(gen.parameters()), allow_unused=True)
    fx["conditional"] = dict(g_init=g_init, d_init=dc_init, noises=noises[:2 * (depth + 1)], z=z, labels=labels, mix=mix,
                             fake=fake.detach().clone(), logits_real=logits_real.detach().clone(),
                             d_loss=float(cdl.detach()), g_loss=float(cgl.detach()),
                             d_grad_norms={n: float(g.norm()) for (n, _), g in zip(dis.named_parameters(), cdg)
                                           if g is not None},
                             g_grad_norms={n: float(g.norm()) for (n, _), g in zip(gen.named_parameters(), cgg)
                                           if g is not None})
    out = os.path.join(ROOT, "tests", "golden", name + ".pt")
    torch.save(fx, out)
    print(name, "->", out, "%.1f KB" % (os.path.getsize(out) / 1024), "standard d_loss %.5f; conditional d/g %.5f %.5f"
          % (fx["standard"]["d_loss"], fx["conditional"]["d_loss"], fx["conditional"]["g_loss"]))


def live_full_width(outdir):
    """BASELINE config 1 at the reference's real channel widths (too large to commit): resolution-128 config,
    depth 1 (8x8) and depth 2 (16x16), batch 4.  Used by tests/test_oracle_vs_reference.py in the build container."""
    full = O.Arch(128)
    golden_trajectory("live_full128", full, batch=4, schedule=[(1, 0.5), (1, 0.75), (2, 0.5), (2, 1.0)],
                      outdir=outdir)


def main():
    if len(sys.argv) >= 2 and sys.argv[1] == "--losses":
        tiny16 = O.Arch(16, latent_size=32, dlatent_size=512, mapping_fmaps=32, mapping_layers=4, fmap_base=128,
                        fmap_max=32)
        golden_standard_and_conditional("losses2_tiny16_d2", tiny16, batch=4, depth=2, alpha=0.5)
        return
    if len(sys.argv) >= 2 and sys.argv[1] == "--full":
        # BASELINE configs 2 and 3 at full width (minutes of CPU): python oracle/make_golden.py --full
        golden_full_width("full128_d5", 128, batch=4, seed=11)
        golden_full_width("full256_d6", 256, batch=4, seed=12)
        return
    if len(sys.argv) >= 3 and sys.argv[1] == "--live":
        os.makedirs(sys.argv[2], exist_ok=True)
        return live_full_width(sys.argv[2])
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    # low-resolution (unfused up/down) trajectory: 10 G+D steps, BASELINE config 1 in miniature
    tiny16 = O.Arch(16, latent_size=32, dlatent_size=512, mapping_fmaps=32, mapping_layers=4, fmap_base=128,
                    fmap_max=32)   # dlatent must stay 512: the reference does not forward it to GSynthesis (GAN.py:244)
    golden_trajectory("traj_tiny16", tiny16, batch=4,
                      schedule=[(1, 0.25), (1, 0.5), (1, 1.0)] + [(2, 0.1 * i) for i in range(1, 8)])
    # depth 0 (no fade branch) forward/backward
    golden_forward_backward("fb_tiny16_d0", tiny16, batch=4, depth=0, alpha=1.0)
    # fused up-conv (64->128) and fused down-conv (128->64): resolution 128 with thin channels
    tiny128 = O.Arch(128, latent_size=32, dlatent_size=512, mapping_fmaps=32, mapping_layers=2, fmap_base=256,
                     fmap_max=16)
    golden_forward_backward("fb_tiny128_d5", tiny128, batch=4, depth=5, alpha=0.5)
    golden_forward_backward("fb_tiny128_d4", tiny128, batch=4, depth=4, alpha=0.3, seed=2)
    # the other loss classes the reference ships (configs: loss = 'hinge' | 'relativistic-hinge')
    golden_other_losses("losses_tiny16_d2", tiny16, batch=4, depth=2, alpha=0.5)


if __name__ == "__main__":
    main()
