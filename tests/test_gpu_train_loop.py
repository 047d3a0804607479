"""StyleGAN.train() on the GPU (reference models/GAN.py:682-826, train.py:102-126): the progressive loop over
two depths on a 64-image synthetic dataset, its five checkpoint files per epoch plus the extra GAN_STATE file, and
`resume=` after a simulated crash -- the resumed run must end where the uninterrupted one ends."""
import logging
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

import stylegan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture
def data_module(monkeypatch):
    """`train()` imports the reference's `data` package (out of scope: reused from the reference unchanged).  On the GPU
    box the reference is absent, so the one function the loop needs is supplied with the reference's settings
    (data/__init__.py:32-52: shuffle, drop_last, pin_memory)."""
    mod = types.ModuleType("data")

    def get_data_loader(dataset, batch_size, num_workers):
        from torch.utils.data import DataLoader
        return DataLoader(dataset, batch_size=batch_size, shuffle=True, num_workers=num_workers, drop_last=True,
                          pin_memory=True)
    mod.get_data_loader = get_data_loader
    monkeypatch.setitem(sys.modules, "data", mod)
    return mod


class _Images(torch.utils.data.Dataset):
    def __init__(self, n, res):
        g = torch.Generator().manual_seed(123)
        self.x = torch.randn(n, 3, res, res, generator=g).clamp_(-1, 1)

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i]


class _Crash(Exception):
    pass


class _Logger(logging.Logger):
    """stops the run when a message starts with `stop_at` (a simulated crash right after a checkpoint was written)"""

    def __init__(self, stop_at=None):
        super().__init__("sgtest")
        self.stop_at, self.lines = stop_at, []

    def info(self, msg, *a, **k):
        line = msg % a if a else msg
        self.lines.append(line)
        if self.stop_at is not None and line.startswith(self.stop_at) and \
                any(l.startswith("Currently working on depth: 3") for l in self.lines):
            raise _Crash()


def _make(arch, g_sd, d_sd):
    import models
    from models import _ops
    from models.GAN import StyleGAN
    models.set_precision("fp32")
    gk, dk = arch.g_kwargs(), arch.d_kwargs()
    for k in ("resolution", "num_channels", "structure"):
        gk.pop(k), dk.pop(k)
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)
    sg = StyleGAN("linear", arch.resolution, 3, arch.latent_size, gk, dk, opt, dict(opt), loss="logistic", use_ema=True,
                  device=torch.device(DEV))
    sg.gen.load_state_dict(g_sd); sg.dis.load_state_dict(d_sd); sg.gen_shadow.load_state_dict(g_sd)
    _ops.clear_weight_cache()
    return sg


def _seed(s):
    from models import CustomLayers
    torch.manual_seed(s); random.seed(s); np.random.seed(s); CustomLayers.manual_seed(s)


def test_train_loop_checkpoint_and_resume(tmp_path, data_module, monkeypatch):
    from models.GAN import StyleGAN
    monkeypatch.setattr(StyleGAN, "create_grid", staticmethod(lambda samples, scale_factor, img_file: None))
    arch = O.Arch(16, fmap_base=512, fmap_max=32, latent_size=32, mapping_fmaps=32)
    gen = torch.Generator().manual_seed(3)
    g_sd, d_sd = O.init_generator(arch, gen), O.init_discriminator(arch, gen)
    ds = _Images(64, 16)
    sched = dict(num_workers=0, epochs=[1, 2, 2], batch_sizes=[8, 8, 8], fade_in_percentage=[50, 50, 50],
                 num_samples=4, start_depth=1, feedback_factor=2, checkpoint_factor=1)

    # ---- run A: uninterrupted over depths 1 and 2 (8x8 and 16x16)
    out_a = str(tmp_path / "a")
    _seed(11)
    a = _make(arch, g_sd, d_sd)
    log_a = _Logger()
    a.train(dataset=ds, logger=log_a, output=out_a, **sched)
    files = sorted(os.listdir(os.path.join(out_a, "models")))
    for d, e in ((1, 1), (1, 2), (2, 1), (2, 2)):
        for stem in ("GAN_GEN", "GAN_DIS", "GAN_GEN_OPTIM", "GAN_DIS_OPTIM", "GAN_GEN_SHADOW", "GAN_STATE"):
            assert "%s_%d_%d.pth" % (stem, d, e) in files, (stem, d, e)
    assert any(l.startswith("Elapsed:") and "D_Loss" in l for l in log_a.lines)      # the reference's progress line
    assert a._graphs.entries and any(e.graph is not None for e in a._graphs.entries.values())   # replayed, not eager

    # ---- run B: same seeds, crashes at the start of epoch 2 of the last depth (the epoch-1 checkpoint exists)
    out_b = str(tmp_path / "b")
    _seed(11)
    b = _make(arch, g_sd, d_sd)
    with pytest.raises(_Crash):
        b.train(dataset=ds, logger=_Logger(stop_at="Epoch: [2]"), output=out_b, **sched)
    del b
    mdir = os.path.join(out_b, "models")
    assert os.path.exists(os.path.join(mdir, "GAN_STATE_2_1.pth")) and not os.path.exists(os.path.join(mdir, "GAN_GEN_2_2.pth"))

    # ---- run C: a fresh process would do exactly this (reference train.py:102-126), then resume
    _seed(999)                                               # the restored RNG state must win over whatever is current
    c = _make(arch, g_sd, d_sd)
    c.gen.load_state_dict(torch.load(os.path.join(mdir, "GAN_GEN_2_1.pth")))
    c.dis.load_state_dict(torch.load(os.path.join(mdir, "GAN_DIS_2_1.pth")))
    c.gen_shadow.load_state_dict(torch.load(os.path.join(mdir, "GAN_GEN_SHADOW_2_1.pth")))
    c.gen_optim.load_state_dict(torch.load(os.path.join(mdir, "GAN_GEN_OPTIM_2_1.pth")))
    c.dis_optim.load_state_dict(torch.load(os.path.join(mdir, "GAN_DIS_OPTIM_2_1.pth")))
    from models import _ops
    _ops.clear_weight_cache()
    log_c = _Logger()
    c.train(dataset=ds, logger=log_c, output=out_b, resume=os.path.join(mdir, "GAN_STATE_2_1.pth"), **sched)
    assert any(l.startswith("Resuming at depth 3, epoch 2") for l in log_c.lines)
    assert os.path.exists(os.path.join(mdir, "GAN_GEN_2_2.pth"))

    # the resumed run ends exactly where the uninterrupted one ends: same data order, latents, mixing draws, noise
    # stream and fade-in position, and -- every reduction being order-fixed and the replayed step executing the same
    # arithmetic as the eager one -- bitwise the same parameters
    last_a = [l for l in log_a.lines if l.startswith("Elapsed:")][-1].split("Step:")[1]
    last_c = [l for l in log_c.lines if l.startswith("Elapsed:")][-1].split("Step:")[1]
    assert last_a == last_c, (last_a, last_c)                                          # step, batch, both losses
    for name, ma, mc in (("gen", a.gen, c.gen), ("dis", a.dis, c.dis), ("shadow", a.gen_shadow, c.gen_shadow)):
        sa, sc = ma.state_dict(), mc.state_dict()
        for k in sa:
            if name == "shadow" and k == "truncation.avg_latent":
                # the shadow's W average moves only when sample grids are rendered from `fixed_input`, which train()
                # draws afresh at every call (reference GAN.py:718): not part of the resumable state
                continue
            assert torch.equal(sc[k], sa[k]), (name, k)
