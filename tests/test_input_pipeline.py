"""GPU input pipeline (models/_input.py): host-side pieces on CPU; the kernel itself in the `gpu` tests below."""
import numpy as np
import pytest
import torch

from models import _input


def test_collate_and_flip_mask_follow_torchvision_rng():
    imgs = [np.random.RandomState(i).randint(0, 256, (6, 5, 3), dtype=np.uint8) for i in range(4)]
    b = _input.collate_uint8(imgs)
    assert b.dtype == torch.uint8 and b.shape == (4, 6, 5, 3) and np.array_equal(b[2].numpy(), imgs[2])
    with pytest.raises(TypeError):
        _input.collate_uint8([np.zeros((2, 2, 3), dtype=np.float32)])
    torch.manual_seed(7)
    want = [bool(torch.rand(1) < 0.5) for _ in range(8)]        # what RandomHorizontalFlip would decide, image by image
    torch.manual_seed(7)
    assert _input.flip_mask(8).bool().tolist() == want


def test_reference_transform_matches_torchvision_formula():
    u8 = torch.randint(0, 256, (3, 4, 5, 3), dtype=torch.uint8)
    flip = torch.tensor([1, 0, 1], dtype=torch.uint8)
    got = _input.reference_transform(u8, flip)
    for n in range(3):
        img = u8[n].flip(1) if flip[n] else u8[n]
        t = img.permute(2, 0, 1).float().div(255)                 # ToTensor
        t = (t - 0.5) / 0.5                                       # Normalize(0.5, 0.5)
        assert torch.equal(got[n], t)
    assert float(got.min()) >= -1.0 and float(got.max()) <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 16, 16, 3), (3, 7, 5, 3), (2, 9, 4, 1), (1, 256, 256, 3)])
def test_prep_images_bit_exact_on_gpu(shape):
    torch.manual_seed(3)
    u8 = torch.randint(0, 256, shape, dtype=torch.uint8)
    flip = _input.flip_mask(shape[0])
    out = _input.prep_images(u8, flip)
    assert out.is_cuda and out.dtype == torch.float32 and out.shape == (shape[0], shape[3], shape[1], shape[2])
    assert torch.equal(out.cpu(), _input.reference_transform(u8, flip))
    assert torch.equal(_input.prep_images(u8).cpu(), _input.reference_transform(u8))
