"""Host logic of the resumable / data-parallel training loop, no GPU: schedule arithmetic and RNG
round trip of models/_train_state.py, shard disjointness of models/_ddp.sharded_loader."""
import random

import numpy as np
import torch

from models import _ddp, _train_state


def test_resume_point_schedule():
    epochs = [2, 3, 4]
    st = _train_state.capture(depth=1, epoch=2, ticker=37, step=101)
    assert _train_state.resume_point(st, epochs) == (1, 3, 37, 101)          # mid-depth: next epoch, same ticker
    st = _train_state.capture(depth=1, epoch=3, ticker=55, step=150)
    assert _train_state.resume_point(st, epochs) == (2, 1, 1, 150)           # depth finished: next depth from alpha 0
    st = _train_state.capture(depth=2, epoch=4, ticker=9, step=400)
    assert _train_state.resume_point(st, epochs)[0] == 3                     # schedule finished: nothing left to run
    bad = dict(st, version=99)
    try:
        _train_state.resume_point(bad, epochs)
        assert False
    except ValueError:
        pass


def test_rng_round_trip_and_rank_streams(tmp_path):
    from models import CustomLayers
    torch.manual_seed(5); random.seed(5); np.random.seed(5)
    CustomLayers.manual_seed(77)
    CustomLayers._Noise.offset = 1234
    st = _train_state.capture(0, 1, 10, 10)
    path = tmp_path / "GAN_STATE_0_1.pth"
    torch.save(st, path)
    expect = (torch.randn(4), random.random(), random.randint(1, 9), np.random.rand())
    torch.manual_seed(1); random.seed(1); np.random.seed(1); CustomLayers.manual_seed(1)
    loaded = _train_state.load(str(path))
    _train_state.restore_rng(loaded, rank=0)
    got = (torch.randn(4), random.random(), random.randint(1, 9), np.random.rand())
    assert torch.equal(got[0], expect[0]) and got[1:] == expect[1:]
    assert (CustomLayers._Noise.seed, CustomLayers._Noise.offset) == (77, 1234)
    # other ranks: deterministic, different from rank 0 and from each other
    _train_state.restore_rng(loaded, rank=1)
    a = torch.randn(4)
    _train_state.restore_rng(loaded, rank=2)
    b = torch.randn(4)
    _train_state.restore_rng(loaded, rank=1)
    assert torch.equal(torch.randn(4), a) and not torch.equal(a, b) and not torch.equal(a, expect[0])


def test_sharded_loader_partitions_every_epoch():
    data = torch.arange(103).float().unsqueeze(1)
    ds = torch.utils.data.TensorDataset(data)
    world, bs = 4, 5
    loaders = [_ddp.sharded_loader(ds, bs, 0, r, world) for r in range(world)]
    orders = []
    for epoch in (0, 1):
        seen = []
        for ld in loaders:
            ld.sampler.set_epoch(epoch)
            items = [int(v) for (b,) in ld for v in b.flatten()]
            assert len(items) == len(ld) * bs and len(ld) == (103 // world) // bs      # equal, full batches only
            seen.append(items)
        flat = [v for s in seen for v in s]
        assert len(set(flat)) == len(flat)                                           # disjoint shards
        orders.append(seen[0])
    assert orders[0] != orders[1]                                                    # reshuffled per epoch


class _Items(torch.utils.data.Dataset):
    def __init__(self, n):
        self.x = torch.arange(n).float().view(n, 1, 1, 1).expand(n, 3, 4, 4).contiguous()

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i]


def _fake_gan(calls):
    """A StyleGAN whose networks and update steps are stubs: only the control flow of train() is exercised."""
    import torch.nn as nn
    from models.GAN import StyleGAN

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.zeros(1))

        def forward(self, z, depth, alpha, labels_in=None):
            return torch.zeros(z.shape[0], 3, 4, 4)

    sg = StyleGAN.__new__(StyleGAN)
    sg.depth, sg.structure, sg.latent_size, sg.use_ema, sg.conditional = 3, "linear", 8, True, False
    sg.device, sg.ddp = torch.device("cpu"), None
    sg.gen, sg.dis, sg.gen_shadow = Net(), Net(), Net()
    sg.gen_optim = torch.optim.SGD(sg.gen.parameters(), lr=0.1)
    sg.dis_optim = torch.optim.SGD(sg.dis.parameters(), lr=0.1)

    def d_step(z, img, depth, alpha, labels=None):
        calls.append(("d", depth, round(float(alpha), 9), round(float(z.sum()), 5), round(float(img.sum()), 3)))
        return 0.5

    def g_step(z, img, depth, alpha, labels=None):
        calls.append(("g", depth, round(float(alpha), 9)))
        return 0.25

    sg.optimize_discriminator, sg.optimize_generator = d_step, g_step
    sg.create_grid = lambda **kw: None
    return sg


def test_train_loop_resume_continues_the_schedule(tmp_path, monkeypatch):
    import logging
    import sys
    import types
    stub = types.ModuleType("data")

    def get_data_loader(dataset, batch_size, num_workers):
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=True, drop_last=True)

    stub.get_data_loader = get_data_loader
    monkeypatch.setitem(sys.modules, "data", stub)
    logger = logging.getLogger("test_train_state")
    kw = dict(dataset=_Items(18), num_workers=0, epochs=[2, 3, 2], batch_sizes=[4, 4, 8], fade_in_percentage=[50, 50, 50],
              logger=logger, num_samples=4, feedback_factor=2)

    torch.manual_seed(3); random.seed(3); np.random.seed(3)
    full = []
    out_a = tmp_path / "a"
    _fake_gan(full).train(output=str(out_a), **kw)
    files = sorted(p.name for p in (out_a / "models").iterdir())
    assert "GAN_GEN_1_2.pth" in files and "GAN_STATE_1_2.pth" in files and "GAN_GEN_SHADOW_2_2.pth" in files
    assert len([f for f in files if f.startswith("GAN_STATE_")]) == 2 + 3 + 2
    # alpha ramps from 1/fade_point to 1 over the first half of each depth and stays there
    d1 = [c for c in full if c[0] == "d" and c[1] == 1]
    assert len(d1) == 3 * 4 and d1[0][2] == round(1 / 6, 9) and d1[5][2] == 1.0 and d1[-1][2] == 1.0

    # a second process resumes after depth 1, epoch 1 with different ambient RNG: identical remaining calls
    torch.manual_seed(99); random.seed(99); np.random.seed(99)
    tail = []
    _fake_gan(tail).train(output=str(tmp_path / "b"), resume=str(out_a / "models" / "GAN_STATE_1_1.pth"), **kw)
    first = next(i for i, c in enumerate(full) if c[0] == "d" and c[1] == 1) + 2 * 4      # skip depth 1, epoch 1
    assert tail == full[first:]
    # resuming from the last epoch of a depth starts the next depth at alpha = 1 / fade_point
    tail2 = []
    _fake_gan(tail2).train(output=str(tmp_path / "c"), resume=str(out_a / "models" / "GAN_STATE_1_3.pth"), **kw)
    assert tail2 == [c for c in full if c[1] == 2]
