"""GPU input pipeline: ship decoded uint8 images to the device and do the reference's per-image
transform chain there.

Reference: `data/transforms.py:11-33` = RandomHorizontalFlip -> [Resize] -> ToTensor -> Normalize(0.5, 0.5) inside the
DataLoader workers, float32 CHW tensors through pinned memory (`data/__init__.py:32-52`).  Here the workers only decode
(and resize) to uint8 HWC; `collate_uint8` stacks them into one pinned uint8 batch (4x fewer bytes over PCIe), and
`prep_images` produces the float32 NCHW batch `StyleGAN.optimize_*` expects with one kernel (`sg_image_u8_to_f32`),
bit-identical to the CPU transform.  The flip decisions are drawn on the host generator like torchvision does
(`torch.rand(1) < p` per image), so a seeded run flips the same images.  Not yet run on hardware."""
import numpy as np
import torch

from ._lib import lib, ptr, stream


def collate_uint8(samples):
    """DataLoader collate_fn: list of HWC uint8 arrays / PIL images / tensors -> pinned uint8 [B,H,W,C]."""
    arrs = []
    for s in samples:
        a = s if isinstance(s, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(s)))
        if a.dtype != torch.uint8:
            raise TypeError("collate_uint8 expects decoded uint8 images, got %s" % a.dtype)
        if a.dim() == 2:
            a = a.unsqueeze(-1)
        arrs.append(a)
    out = torch.stack(arrs, 0).contiguous()
    return out.pin_memory() if torch.cuda.is_available() else out


def flip_mask(n, p=0.5):
    """One RandomHorizontalFlip decision per image, drawn like torchvision (`torch.rand(1) < p`, host generator)."""
    return torch.tensor([bool(torch.rand(1) < p) for _ in range(n)], dtype=torch.uint8)


def prep_images(u8, flip=None, device=None):
    """u8 [B,H,W,C] uint8 (host or device), flip [B] uint8/bool or None -> float32 [B,C,H,W] in [-1, 1] on the device."""
    if u8.dtype != torch.uint8 or u8.dim() != 4:
        raise TypeError("prep_images expects a uint8 [B,H,W,C] batch")
    if device is None:
        device = u8.device if u8.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = u8.to(device, non_blocking=True).contiguous()
    f = None
    if flip is not None:
        f = flip.to(torch.uint8).to(device, non_blocking=True).contiguous()
        if f.numel() != x.shape[0]:
            raise ValueError("flip mask must have one entry per image")
    B, H, W, C = x.shape
    if not 1 <= C <= 4:
        raise ValueError("prep_images supports 1..4 channels")
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=device)
    lib.sg_image_u8_to_f32(ptr(x), ptr(f), ptr(out), B, H, W, C, stream())
    return out


def reference_transform(u8, flip=None):
    """The same arithmetic with torch ops on the host (ToTensor's .div(255), Normalize's .sub_(0.5).div_(0.5)): the
    checker for prep_images."""
    x = u8.cpu()
    if flip is not None:
        m = flip.cpu().bool()
        x = torch.where(m.view(-1, 1, 1, 1), x.flip(2), x)
    t = x.permute(0, 3, 1, 2).to(torch.float32).div(255)
    return t.sub_(0.5).div_(0.5)
