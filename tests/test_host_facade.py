"""models/_host_facade.py: CPU-resident module + CPU tensors -> executed on a device replica, CPU results (the calling
convention of the reference's generate_* scripts).  The control flow is checked on CPU with the device transfer stubbed;
the end-to-end case is a `gpu` test."""
import pytest
import torch
import torch.nn as nn

from models import _host_facade as hf


class _M(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.ones(3))
        self.register_buffer("avg", torch.zeros(3))

    def forward(self, x, k=1.0):
        if self.training:
            self.avg.copy_(self.avg * 0.5 + x[0] * 0.5)
        return x * self.w * k


def test_facade_control_flow_with_stubbed_device(monkeypatch):
    monkeypatch.setattr(hf.torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(hf.torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(nn.Module, "to", lambda self, *a, **k: self)
    monkeypatch.setattr(hf, "_to_device", lambda x, dev: x)
    m = _M()
    x = torch.arange(6.).view(2, 3)
    assert hf.wants_facade(m, x) and not hf.wants_facade(m, None)
    y = hf.call(m, x, k=2.0)
    assert torch.equal(y, x * 2) and torch.equal(m.avg, x[0] * 0.5)          # buffer update copied back
    rep = m.__dict__[hf._KEY][1]
    assert rep is not m and hf.call(m, x) is not None
    assert m.__dict__[hf._KEY][1] is rep and torch.allclose(m.avg, x[0] * 0.75)   # replica reused
    with torch.no_grad():
        m.w.mul_(3)                                                            # e.g. load_state_dict
    assert torch.equal(hf.call(m, x), x * 3) and m.__dict__[hf._KEY][1] is not rep
    m.eval()
    before = m.avg.clone()
    hf.call(m, x)
    assert torch.equal(m.avg, before)


def test_no_facade_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    assert not hf.wants_facade(_M(), torch.zeros(2, 3))


@pytest.mark.gpu
def test_generate_script_calling_convention_on_gpu():
    """CPU Generator + CPU latents, as generate_samples.py / generate_mixing_figure.py use it."""
    import numpy as np
    from models.GAN import Generator
    torch.manual_seed(0)
    gen = Generator(resolution=32, latent_size=32, dlatent_size=512, mapping_fmaps=32, mapping_layers=2, fmap_base=256,
                    fmap_max=32)
    ref = Generator(resolution=32, latent_size=32, dlatent_size=512, mapping_fmaps=32, mapping_layers=2, fmap_base=256,
                    fmap_max=32)
    ref.load_state_dict(gen.state_dict())
    ref = ref.cuda().eval()
    gen.eval()
    z = torch.randn(3, 32)
    with torch.no_grad():
        img = gen(z, depth=3, alpha=1)
        dl = gen.g_mapping(z)
        img2 = gen.g_synthesis(dl, depth=3, alpha=1)
        want = ref(z.cuda(), 3, 1)
    assert img.device.type == "cpu" and img.shape == (3, 3, 32, 32) and dl.device.type == "cpu"
    assert isinstance(dl.numpy(), np.ndarray) and isinstance(gen.truncation.avg_latent.numpy(), np.ndarray)
    assert torch.allclose(img, want.cpu(), atol=1e-5) and torch.allclose(img2, img, atol=1e-5)
    # training mode (the scripts never call eval()): the W average moves and comes back to the host module
    gen.train()
    torch.manual_seed(1)
    before = gen.truncation.avg_latent.clone()
    with torch.no_grad():
        gen(z, depth=3, alpha=1)
    assert not torch.equal(gen.truncation.avg_latent, before) and gen.truncation.avg_latent.device.type == "cpu"
