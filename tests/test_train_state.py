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
