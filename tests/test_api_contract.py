"""CPU-side checks of the drop-in boundary: C-ABI exports, state_dict key/shape/order contract, loud failure off-GPU."""
import ctypes
import os

import pytest
import torch

import stylegan_oracle as O
from test_oracle_golden import arch_from

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nets(arch):
    from models.GAN import Generator, Discriminator
    gk, dk = arch.g_kwargs(), arch.d_kwargs()
    return Generator(**gk), Discriminator(**dk)


def test_library_exports_every_declared_symbol():
    from models import _lib
    sigs = _lib.parse_header()
    assert len(sigs) >= 45
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in sigs:
        assert hasattr(dll, name), "libsgb200.so does not export %s" % name
    assert dll.sg_abi_version() == 2
    assert dll.sg_geom_size() == ctypes.sizeof(_lib.SgConvGeom)
    assert dll.sg_entry_size() == 64


@pytest.mark.parametrize("name", ["traj_tiny16", "fb_tiny128_d5"])
def test_state_dict_contract_matches_reference(golden_dir, name):
    """Key names, order and shapes equal those recorded from the reference modules (checkpoint compatibility,
    and module.parameters() order for the optimizer files)."""
    fx = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    g, d = _nets(arch_from(fx["arch"]))
    for net, ref in ((g, fx["g_init"]), (d, fx["d_init"])):
        sd = net.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in ref:
            assert tuple(sd[k].shape) == tuple(ref[k].shape), k
        net.load_state_dict(ref)          # strict
    assert [n for n, _ in g.named_parameters()] == O.param_names(fx["g_init"])
    assert [n for n, _ in d.named_parameters()] == O.param_names(fx["d_init"])


def test_full_size_key_counts():
    from models.GAN import Generator, Discriminator
    for res, ng, nd in ((128, 86, 48), (256, 99, 56)):
        g = Generator(resolution=res, latent_size=512, mapping_layers=4, blur_filter=[1, 2, 1], truncation_psi=0.7,
                      truncation_cutoff=8, num_channels=3, structure="linear", unknown_kwarg=1)
        d = Discriminator(resolution=res, num_channels=3, use_wscale=True, blur_filter=[1, 2, 1],
                          structure="linear", unknown_kwarg=1)
        assert len(g.state_dict()) == ng and len(d.state_dict()) == nd
    g = Generator(resolution=64, truncation_psi=-1.0)
    assert g.truncation is None and "truncation.avg_latent" not in g.state_dict()


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: the product path refuses CPU tensors instead of silently running torch."""
    from models.GAN import Generator, StyleGAN
    if torch.cuda.is_available():
        pytest.skip("with a GPU present CPU-resident generators are served by models/_host_facade.py")
    g = Generator(resolution=8, latent_size=16, dlatent_size=512, mapping_fmaps=16, mapping_layers=2, fmap_max=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        g(torch.randn(2, 16), 0, 1.0)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)
    with pytest.raises(RuntimeError, match="CUDA"):
        StyleGAN("linear", 8, 3, 16, dict(latent_size=16), dict(), opt, opt, loss="logistic", device="cpu")


def test_reference_error_conventions():
    from models.GAN import Generator, Discriminator
    d = Discriminator(resolution=8, fmap_max=8)
    with pytest.raises(AssertionError):
        d(torch.zeros(1, 3, 64, 64), depth=5)
    with pytest.raises(NotImplementedError):
        Discriminator(resolution=8, nonlinearity="relu")


def test_fused_adam_state_dict_layout():
    from models._optim import FusedAdam
    ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
    opt = FusedAdam(ps, lr=0.003, betas=(0, 0.99), eps=1e-8)
    ref = torch.optim.Adam(ps, lr=0.003, betas=(0.0, 0.99), eps=1e-8)
    a, b = opt.state_dict(), ref.state_dict()
    assert a["param_groups"][0]["params"] == b["param_groups"][0]["params"]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert a["param_groups"][0][k] == b["param_groups"][0][k]
    # a reference optimizer checkpoint loads into the fused optimizer
    for p in ps:
        p.grad = torch.ones_like(p)
    ref.step()
    opt.load_state_dict(ref.state_dict())
    st = opt.state[ps[0]]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 1.0


def test_bucketize():
    from models._ddp import bucketize
    assert bucketize([5, 5, 5, 20, 1], 10) == [[0, 1], [2], [3], [4]]
    assert bucketize([], 10) == []


def test_weight_cache_is_keyed_on_the_parameter_object():
    """ids and storage pointers are recycled after a model is freed: a fresh Parameter must never hit the derived
    copy of a dead one (models/_ops.py cached_weight).  Variants live in persistent buffers that are
    rebuilt IN PLACE when the parameter's update counter moves."""
    import gc
    from models import _ops
    _ops.clear_weight_cache()

    def builder(_):
        def build(W, out):
            if out is None:
                return W.detach().clone()
            out.copy_(W.detach())
            return out
        return build
    seen = []
    for i in range(1, 33):
        W = torch.nn.Parameter(torch.full((8, 8), float(i)))
        E = _ops.cached_weight(W, "k", builder(W))
        again = _ops.cached_weight(W, "k", lambda W, out: (_ for _ in ()).throw(AssertionError("must be a cache hit")))
        assert again is E
        seen.append(float(E[0, 0]))
        del W, E, again
        gc.collect()
    assert seen == [float(i) for i in range(1, 33)]
    assert len(_ops._prep_cache) == 0           # entries die with their parameters
    # an edit through autograd-visible ops (version bump) makes a new entry; a raw-pointer rewrite announced through the
    # update counters (optimizer token / global epoch) refreshes the SAME buffer in place
    W = torch.nn.Parameter(torch.zeros(4))
    a = _ops.cached_weight(W, "k", builder(W))
    with torch.no_grad():
        W.add_(1.0)
    b = _ops.cached_weight(W, "k", builder(W))
    assert float(a[0]) == 0.0 and float(b[0]) == 1.0 and b is not a
    W.data.fill_(5.0)
    _ops.bump_weights_epoch()
    c = _ops.cached_weight(W, "k", builder(W))
    assert c is b and float(c[0]) == 5.0
    W.data.fill_(7.0)
    W._sg_token = [0]
    d = _ops.cached_weight(W, "k", builder(W))           # new owner token -> stale -> rebuilt in place
    assert d is b and float(d[0]) == 7.0
    W.data.fill_(9.0)
    W._sg_token[0] += 1
    _ops.refresh_derived([W])
    assert float(b[0]) == 9.0
    assert _ops.cached_weight(W, "k", lambda W, out: (_ for _ in ()).throw(AssertionError("current: no rebuild"))) is b


def test_refresh_derived_batches_recorded_weight_preparations(monkeypatch):
    """refresh_derived sends every variant that was built by one recorded sg_wprep_fwd into a bf16 buffer to ONE
    sg_wprep_fwd_multi launch and rebuilds the others through their build closures (host logic; the kernel itself is
    tested on the GPU: tests/test_gpu_fast_paths.py)."""
    from models import _ops
    _ops.clear_weight_cache()
    W1 = torch.nn.Parameter(torch.zeros(4, 4, 3, 3))
    W2 = torch.nn.Parameter(torch.zeros(2, 2))
    tok = [0]
    W1._sg_token = W2._sg_token = tok
    coef = torch.zeros(9, 9)
    rebuilt = []

    def recorded_build(W, out):                       # what _ops.wprep_fwd leaves behind for cached_weight
        E = out if out is not None else torch.zeros(9, 4, 4, dtype=torch.bfloat16)
        rebuilt.append("recorded")
        _ops._last_wprep[0] = (E.data_ptr(), (W.data_ptr(), E.data_ptr(), coef.data_ptr(), 4, 4, 9, 9, 1, 0.5, coef))
        return E

    def plain_build(W, out):
        rebuilt.append("plain")
        return out if out is not None else W.detach().clone()

    E1 = _ops.cached_weight(W1, "tc", recorded_build)
    E2 = _ops.cached_weight(W2, "other", plain_build)
    assert rebuilt == ["recorded", "plain"]
    batches = []
    monkeypatch.setattr(_ops, "_wprep_multi", lambda descs, dev: batches.append([d[:9] for d in descs]))
    tok[0] += 1
    _ops.refresh_derived([W1, W2])
    assert rebuilt == ["recorded", "plain", "plain"]   # W2's variant through its closure, W1's through the batch
    assert batches == [[(W1.data_ptr(), E1.data_ptr(), coef.data_ptr(), 4, 4, 9, 9, 1, 0.5)]]
    # both are current now: no rebuild on use
    boom = lambda W, out: (_ for _ in ()).throw(AssertionError("current: no rebuild"))
    assert _ops.cached_weight(W1, "tc", boom) is E1 and _ops.cached_weight(W2, "other", boom) is E2
    _ops.clear_weight_cache()
