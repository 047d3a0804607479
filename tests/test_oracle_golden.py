"""The oracle replayed against golden vectors produced by the live reference (oracle/make_golden.py)."""
import os

import pytest
import torch

import stylegan_oracle as O

TOL = 1e-4  # fp32 CPU vs fp32 CPU, same torch build: differences are summation-order only


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def arch_from(d):
    a = O.Arch(d["resolution"])
    a.__dict__.update(d)
    return a


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", ["fb_tiny16_d0", "fb_tiny128_d5", "fb_tiny128_d4"])
def test_forward_backward(golden_dir, name):
    fx = load(golden_dir, name)
    arch = arch_from(fx["arch"])
    depth, alpha = fx["depth"], fx["alpha"]
    g = {k: v.clone() for k, v in fx["g_init"].items()}
    d = {k: v.clone() for k, v in fx["d_init"].items()}
    gn, dn = O.param_names(g), O.param_names(d)
    for k in gn:
        g[k].requires_grad_(True)
    for k in dn:
        d[k].requires_grad_(True)
    fake = O.generator_forward(g, fx["z"], depth, alpha, arch, z2=fx["mix"][0], mixing_cutoff=fx["mix"][1],
                               noises=fx["noises"])
    assert rel(fake, fx["fake"]) < TOL
    assert rel(g["truncation.avg_latent"], fx["avg_after"]) < TOL
    assert rel(O.discriminator(d, fake, depth, alpha, arch), fx["logits_fake"]) < TOL
    assert rel(O.discriminator(d, fx["real"], depth, alpha, arch), fx["logits_real"]) < TOL
    r1 = O.r1_penalty(d, fx["real"], depth, alpha, arch)
    assert abs(float(r1) - fx["r1"]) <= TOL * abs(fx["r1"]) + 1e-9
    d_loss = O.d_loss_logistic(d, fx["real"], fake.detach(), depth, alpha, arch)
    assert abs(float(d_loss) - fx["d_loss"]) <= TOL * abs(fx["d_loss"])
    grads = torch.autograd.grad(d_loss, [d[k] for k in dn], allow_unused=True)
    for k, gr in zip(dn, grads):
        if gr is None:
            assert k not in fx["d_grads"]
            continue
        ref = fx["d_grads"][k]
        assert abs(float(gr.norm()) - ref["norm"]) <= 1e-3 * ref["norm"] + 1e-9, k   # cancellation-heavy grads (bias before InstanceNorm)
        if ref["full"] is not None:
            assert rel(gr, ref["full"]) < 1e-3, k
    g_loss = O.g_loss_logistic(d, fake, depth, alpha, arch)
    assert abs(float(g_loss) - fx["g_loss"]) <= TOL * abs(fx["g_loss"])
    grads = torch.autograd.grad(g_loss, [g[k] for k in gn], allow_unused=True)
    for k, gr in zip(gn, grads):
        if gr is None:
            assert k not in fx["g_grads"]
            continue
        ref = fx["g_grads"][k]
        assert abs(float(gr.norm()) - ref["norm"]) <= 1e-3 * ref["norm"] + 1e-9, k   # cancellation-heavy grads (bias before InstanceNorm)
        assert rel(gr.flatten()[:64], ref["head"]) < 2e-3, k


def test_trajectory(golden_dir):
    fx = load(golden_dir, "traj_tiny16")
    arch = arch_from(fx["arch"])
    st = O.TrainState(arch, fx["g_init"], fx["d_init"], use_ema=True, ema_decay=0.999)
    for (depth, alpha), s in zip(fx["schedule"], fx["steps"]):
        nz = fx["noises"][:2 * (depth + 1)]
        dl, gl = O.train_step(st, s["images"], s["z"], depth, alpha, s["mix_d"], s["mix_g"], nz, nz)
        assert abs(dl - s["d_loss"]) <= 5e-4 * abs(s["d_loss"]), (depth, alpha)
        assert abs(gl - s["g_loss"]) <= 5e-4 * abs(s["g_loss"]), (depth, alpha)
    # Adam with beta1=0 turns a round-off-level gradient (e.g. init_block.bias of a channel that stays in the linear
    # region of LeakyReLU before InstanceNorm: analytically zero) into a full +-lr step of random sign, so final
    # parameters are only comparable to ~n_steps*lr; the per-step losses above are the tight check.
    PT = 2e-2
    for k, v in fx["g_final"].items():
        assert rel(st.g[k], v) < PT, k
    for k, v in fx["d_final"].items():
        assert rel(st.d[k], v) < PT, k
    for k, v in fx["shadow_final"].items():
        # the reference shadow's avg_latent is its own buffer, only touched when the shadow is run (GAN.py:789)
        if k == "truncation.avg_latent":
            continue
        assert rel(st.shadow[k], v) < PT, k
    for k, v in fx["g_exp_avg_sq"].items():
        assert rel(st.g_opt.v[k], v) < 5e-3, k
        assert st.g_opt.steps[k] == fx["g_adam_step"][k], k
    for k, v in fx["d_exp_avg_sq"].items():
        assert rel(st.d_opt.v[k], v) < 5e-3, k
        assert st.d_opt.steps[k] == fx["d_adam_step"][k], k


@pytest.mark.parametrize("loss_name", ["hinge", "relativistic-hinge"])
def test_other_losses(golden_dir, loss_name):
    """Hinge / relativistic-average hinge of the oracle against the reference's loss classes (models/Losses.py:136-189):
    loss values and the gradient norm of every discriminator parameter."""
    fx = load(golden_dir, "losses_tiny16_d2")
    arch = arch_from(fx["arch"])
    d = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fx["d_init"].items()}
    depth, alpha, real, fake = fx["depth"], fx["alpha"], fx["real"], fx["fake"]
    ref = fx["losses"][loss_name]
    if loss_name == "hinge":
        dl, gl = O.d_loss_hinge(d, real, fake, depth, alpha, arch), O.g_loss_hinge(d, fake, depth, alpha, arch)
    else:
        dl = O.d_loss_rahinge(d, real, fake, depth, alpha, arch)
        gl = O.g_loss_rahinge(d, real, fake, depth, alpha, arch)
    assert abs(float(dl) - ref["d_loss"]) <= TOL * abs(ref["d_loss"])
    assert abs(float(gl) - ref["g_loss"]) <= TOL * abs(ref["g_loss"])
    for loss, norms in ((dl, ref["d_grad_norms"]), (gl, ref["g_grad_norms"])):
        names = [k for k in norms]
        grads = torch.autograd.grad(loss, [d[k] for k in names], allow_unused=True, retain_graph=True)
        for k, g in zip(names, grads):
            got = 0.0 if g is None else float(g.norm())
            assert abs(got - norms[k]) <= 1e-3 * norms[k] + 1e-7, (loss_name, k, got, norms[k])


def test_state_dict_key_contract():
    """86/48 keys at 128, 99/56 at 256, 125/72 at 1024 (avg_latent absent when psi<=0)."""
    for res, ng, nd in ((128, 86, 48), (256, 99, 56), (1024, 125, 72)):
        a = O.Arch(res, fmap_base=64, fmap_max=8, latent_size=8, dlatent_size=8, mapping_layers=4)
        assert len(O.init_generator(a)) == ng
        assert len(O.init_discriminator(a)) == nd


@pytest.mark.parametrize("name,deep", [("full128_d5", True), ("full256_d6", False)])
def test_full_width_fixture_matches_oracle(golden_dir, name, deep):
    """BASELINE configs 2/3 at real channel widths: the seed-derived inputs (O.seeded_case) reproduce, through the oracle,
    what the live reference produced (forward at every depth; for 128x128 also logits, R1 and the discriminator loss)."""
    fx = load(golden_dir, name)
    arch = arch_from(fx["arch"])
    depth, alpha = fx["depth"], fx["alpha"]
    case = O.seeded_case(arch, fx["batch"], depth, fx["seed"])
    g = {k: v.clone() for k, v in case["g_sd"].items()}
    with torch.no_grad():
        for d_, ref in fx["lower"].items():
            avg0 = g["truncation.avg_latent"].clone()
            img = O.generator_forward(g, case["z"], d_, 1.0, arch, z2=case["z2"],
                                      mixing_cutoff=min(case["cutoff"], 2 * (d_ + 1)), noises=case["noises"])
            assert rel(img, ref) < TOL, d_
            g["truncation.avg_latent"].copy_(avg0)
        fake = O.generator_forward(g, case["z"], depth, alpha, arch, z2=case["z2"], mixing_cutoff=case["cutoff"],
                                   noises=case["noises"])
    assert rel(fake, fx["fake"]) < TOL
    assert rel(g["truncation.avg_latent"], fx["avg_after"]) < TOL
    if deep:
        d = case["d_sd"]
        with torch.no_grad():
            assert rel(O.discriminator(d, fx["fake"], depth, alpha, arch), fx["logits_fake"]) < TOL
        r1 = O.r1_penalty(d, case["real"], depth, alpha, arch)
        assert abs(float(r1) - fx["r1"]) <= 1e-3 * abs(fx["r1"])
