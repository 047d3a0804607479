"""CUDA-graph replay of the optimizer steps (models/_graphs.py) against the eager path: same seeds -> same host RNG draws
(style mixing), same Philox noise stream, a different fade-in alpha every step -> the same losses and parameters."""
import random

import pytest
import torch

import stylegan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _run(arch, g_sd, d_sd, precision, graphs, steps, batch, depth, loss="logistic"):
    import models
    from models import _ops, CustomLayers
    from models.GAN import StyleGAN
    models.set_precision(precision)
    _ops.clear_weight_cache()
    gk, dk = arch.g_kwargs(), arch.d_kwargs()
    for k in ("resolution", "num_channels", "structure"):
        gk.pop(k), dk.pop(k)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)
    sg = StyleGAN("linear", arch.resolution, 3, arch.latent_size, gk, dk, opt, dict(opt), loss=loss, use_ema=True,
                  device=torch.device(DEV))
    sg.use_cuda_graphs = graphs
    sg.gen.load_state_dict(g_sd); sg.dis.load_state_dict(d_sd); sg.gen_shadow.load_state_dict(g_sd)
    _ops.clear_weight_cache()
    sg.gen.train(); sg.dis.train(); sg.gen_shadow.train()
    torch.manual_seed(5); random.seed(5); CustomLayers.manual_seed(5)
    data = torch.Generator().manual_seed(9)
    res = 2 ** (depth + 2)
    losses = []
    for i in range(steps):
        z = torch.randn(batch, arch.latent_size, generator=data).to(DEV)
        img = torch.randn(batch, 3, arch.resolution, arch.resolution, generator=data).to(DEV)
        alpha = min(1.0, 0.1 + 0.15 * i)
        dl = sg.optimize_discriminator(z, img, depth, alpha)
        gl = sg.optimize_generator(z, img, depth, alpha)
        losses.append((dl, gl))
    captured = sum(1 for e in sg._graphs.entries.values() if e.graph is not None)
    state = {k: v.detach().clone() for k, v in list(sg.gen.state_dict().items()) + list(sg.dis.state_dict().items())}
    models.set_precision("fp32")
    return losses, state, captured


@pytest.mark.parametrize("precision,loss", [("fp32", "logistic"), ("bf16", "logistic"), ("fp32", "relativistic-hinge")])
def test_graph_replay_equals_eager(precision, loss):
    arch = O.Arch(32, fmap_base=2048, fmap_max=128, latent_size=64, mapping_fmaps=64) if precision == "bf16" else \
        O.Arch(16, fmap_base=512, fmap_max=32, latent_size=32, mapping_fmaps=32)
    gen = torch.Generator().manual_seed(1)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    for k in g_sd:
        if k.endswith("noise.weight"):
            g_sd[k] = torch.randn(g_sd[k].shape, generator=gen)
    depth = arch.depth - 1
    eager, st_e, cap_e = _run(arch, g_sd, d_sd, precision, False, 6, 8, depth, loss)
    graph, st_g, cap_g = _run(arch, g_sd, d_sd, precision, True, 6, 8, depth, loss)
    assert cap_e == 0 and cap_g == 2                      # one graph per step kind
    # fp32: the two paths run the same kernels on the same data; what differs is the atomics' summation order
    # (fp64 InstanceNorm partials on the unfused layers) -> round-off level.  bf16: LeakyReLU mask flips amplify it.
    # Every reduction of the library has a fixed summation order and the replayed step executes the same kernels with
    # the same arguments as the eager one (the per-step scalars just come from device memory): BITWISE equal losses and
    # parameters, with a different alpha at every step and fresh style-mixing draws.
    assert eager == graph, (eager, graph)
    for k in st_e:
        assert torch.equal(st_g[k], st_e[k]), k


def test_training_steps_are_bitwise_reproducible():
    """two eager runs from the same state and seeds: identical losses and parameters (no atomics anywhere)"""
    arch = O.Arch(32, fmap_base=2048, fmap_max=128, latent_size=64, mapping_fmaps=64)
    gen = torch.Generator().manual_seed(4)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    for k in g_sd:
        if k.endswith("noise.weight"):
            g_sd[k] = torch.randn(g_sd[k].shape, generator=gen)
    for precision in ("fp32", "bf16"):
        a, st_a, _ = _run(arch, g_sd, d_sd, precision, False, 3, 8, arch.depth - 1)
        b, st_b, _ = _run(arch, g_sd, d_sd, precision, False, 3, 8, arch.depth - 1)
        assert a == b, (precision, a, b)
        for k in st_a:
            assert torch.equal(st_a[k], st_b[k]), (precision, k)


def test_graph_path_is_skipped_when_not_applicable():
    """evaluation / no-grad calls and small numbers of calls stay on the eager path"""
    arch = O.Arch(16, fmap_base=512, fmap_max=32, latent_size=32, mapping_fmaps=32)
    gen = torch.Generator().manual_seed(2)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    _, _, captured = _run(arch, g_sd, d_sd, "fp32", True, 2, 4, 1)
    assert captured == 0                                   # two warm-up calls per step kind, nothing captured yet
