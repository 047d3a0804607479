"""CPU oracle for the StyleGAN G+D training-step hot path.

TEST INFRASTRUCTURE ONLY.  This file restates, as plain functions over flat
``state_dict``-style dictionaries, the arithmetic that the reference
(huangzh13/StyleGAN.pytorch, mounted at /root/reference) performs inside its
``nn.Module`` classes.  It runs on CPU in fp32 (fp64 on request) and is the
checker for the CUDA engine in ``stylegan.pytorch_b200/``: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it.  The product path never does.

Parity pin: the reference's own tests are empty (test/test_Blocks.py:13-21), so
this oracle is pinned against the *live reference* executed in the build
container (tests/test_oracle_vs_reference.py, skipped where /root/reference is
absent) and against golden vectors generated from the live reference by
``oracle/make_golden.py`` and committed under ``tests/golden/``.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
import random as _pyrandom
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

LRELU_SLOPE = 0.2  # models/GAN.py:68,151,347


# --------------------------------------------------------------------------- #
# architecture bookkeeping
# --------------------------------------------------------------------------- #
def nf(stage: int, fmap_base: int = 8192, fmap_decay: float = 1.0, fmap_max: int = 512) -> int:
    """Feature maps at a stage.  models/GAN.py:138-139, 332-333."""
    return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)


def he_scale(fan_in: int, gain: float, lrmul: float = 1.0) -> float:
    """Runtime weight multiplier with use_wscale=True.  models/CustomLayers.py:84-89,120-127."""
    return gain * fan_in ** (-0.5) * lrmul


class Arch:
    """Static shape description shared by the oracle, the engine and the tests."""

    def __init__(self, resolution: int, latent_size: int = 512, dlatent_size: int = 512,
                 mapping_layers: int = 4, mapping_fmaps: Optional[int] = None, num_channels: int = 3,
                 fmap_base: int = 8192, fmap_decay: float = 1.0, fmap_max: int = 512,
                 truncation_psi: float = 0.7, truncation_cutoff: int = 8, dlatent_avg_beta: float = 0.995,
                 style_mixing_prob: float = 0.9, blur: bool = True, mbstd_group_size: int = 4,
                 structure: str = "linear"):
        self.resolution = resolution
        self.res_log2 = int(math.log2(resolution))
        assert 2 ** self.res_log2 == resolution and resolution >= 4
        self.depth = self.res_log2 - 1                     # models/GAN.py:145,344
        self.num_layers = self.res_log2 * 2 - 2            # models/GAN.py:147,242
        self.latent_size = latent_size
        self.dlatent_size = dlatent_size
        self.mapping_layers = mapping_layers
        self.mapping_fmaps = mapping_fmaps if mapping_fmaps is not None else dlatent_size
        self.num_channels = num_channels
        self.fm = (fmap_base, fmap_decay, fmap_max)
        self.truncation_psi = truncation_psi
        self.truncation_cutoff = truncation_cutoff
        self.dlatent_avg_beta = dlatent_avg_beta
        self.style_mixing_prob = style_mixing_prob
        self.blur = blur
        self.mbstd_group_size = mbstd_group_size
        self.structure = structure

    def nf(self, stage: int) -> int:
        return nf(stage, *self.fm)

    # G block i (0-based, output res 2**(i+3)): in nf(i+1) -> out nf(i+2).  models/GAN.py:161-166
    def g_block_channels(self, i: int) -> Tuple[int, int]:
        return self.nf(i + 1), self.nf(i + 2)

    # D block i (0 = highest res 2**(res_log2-i)): in nf(res-1) -> out nf(res-2).  models/GAN.py:352-356
    def d_block_channels(self, i: int) -> Tuple[int, int]:
        res = self.res_log2 - i
        return self.nf(res - 1), self.nf(res - 2)

    def g_kwargs(self) -> dict:
        """kwargs for reference ``Generator(**kw)`` reproducing this architecture."""
        return dict(resolution=self.resolution, latent_size=self.latent_size, dlatent_size=self.dlatent_size,
                    mapping_layers=self.mapping_layers, mapping_fmaps=self.mapping_fmaps,
                    num_channels=self.num_channels, fmap_base=self.fm[0], fmap_decay=self.fm[1],
                    fmap_max=self.fm[2], truncation_psi=self.truncation_psi,
                    truncation_cutoff=self.truncation_cutoff, dlatent_avg_beta=self.dlatent_avg_beta,
                    style_mixing_prob=self.style_mixing_prob,
                    blur_filter=[1, 2, 1] if self.blur else None, structure=self.structure)

    def d_kwargs(self) -> dict:
        return dict(resolution=self.resolution, num_channels=self.num_channels, fmap_base=self.fm[0],
                    fmap_decay=self.fm[1], fmap_max=self.fm[2], use_wscale=True,
                    mbstd_group_size=self.mbstd_group_size,
                    blur_filter=[1, 2, 1] if self.blur else None, structure=self.structure)


# --------------------------------------------------------------------------- #
# layer primitives
# --------------------------------------------------------------------------- #
def lrelu(x: Tensor) -> Tensor:
    return F.leaky_relu(x, LRELU_SLOPE)


def pixel_norm(x: Tensor, eps: float = 1e-8) -> Tensor:
    """models/CustomLayers.py:22-23."""
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + eps)


def eq_linear(x: Tensor, w: Tensor, b: Optional[Tensor], gain: float, lrmul: float = 1.0) -> Tensor:
    """Equalized-LR dense layer.  models/CustomLayers.py:82-103."""
    w_mul = he_scale(w.shape[1], gain, lrmul)
    y = x @ (w * w_mul).t()
    if b is not None:
        y = y + b * lrmul
    return y


def eq_conv(x: Tensor, w: Tensor, b: Optional[Tensor], gain: float) -> Tensor:
    """Plain equalized-LR conv, 'same' padding.  models/CustomLayers.py:137-141,170-171."""
    k = w.shape[-1]
    w_mul = he_scale(w.shape[1] * k * k, gain)
    return F.conv2d(x, w * w_mul, None if b is None else b, padding=k // 2)


def blur3(x: Tensor) -> Tensor:
    """Depthwise [1,2,1]x[1,2,1]/16, pad 1.  models/CustomLayers.py:251-276."""
    c = x.shape[1]
    k1 = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype, device=x.device)
    k = (k1[:, None] * k1[None, :]) / 16.0
    return F.conv2d(x, k[None, None].expand(c, 1, 3, 3), padding=1, groups=c)


def nearest_up2(x: Tensor) -> Tensor:
    """models/CustomLayers.py:27-36 (factor 2, gain 1) == F.interpolate nearest x2 (models/GAN.py:173)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def avg_pool2(x: Tensor) -> Tensor:
    """2x2 box, stride 2.  models/CustomLayers.py:48-64 (blur kernel [.5,.5]^2, stride 2) and GAN.py:382."""
    return F.avg_pool2d(x, 2)


def up_conv(x: Tensor, w: Tensor, b: Tensor, gain: float, blur: bool = True) -> Tensor:
    """conv0_up of a synthesis block.  models/CustomLayers.py:137-180 with upscale=True, intermediate=Blur.

    out < 128: nearest x2 -> conv3x3 (no bias) -> blur -> +bias.
    out >= 128: transposed conv with the 3x3 kernel box-summed to 4x4 (stride 2, pad 1) -> blur -> +bias
    (CustomLayers.py:143-152).  The two are different functions of the same weights (author's note :144).
    """
    k = w.shape[-1]
    ws = w * he_scale(w.shape[1] * k * k, gain)
    if min(x.shape[2:]) * 2 >= 128:
        wt = F.pad(ws.permute(1, 0, 2, 3), [1, 1, 1, 1])
        wt = wt[:, :, 1:, 1:] + wt[:, :, :-1, 1:] + wt[:, :, 1:, :-1] + wt[:, :, :-1, :-1]
        y = F.conv_transpose2d(x, wt, stride=2, padding=1)
    else:
        y = F.conv2d(nearest_up2(x), ws, None, padding=1)
    if blur:
        y = blur3(y)
    return y + b.view(1, -1, 1, 1)


def down_conv(x: Tensor, w: Tensor, b: Tensor, gain: float) -> Tensor:
    """conv1_down of a discriminator block.  models/CustomLayers.py:156-180 with downscale=True.

    in >= 128: conv with the 3x3 kernel box-averaged to 4x4, stride 2, pad 1, then +bias (:158-165).
    in < 128: conv3x3 (no bias) -> 2x2 box stride 2 -> +bias (:166-180).
    """
    k = w.shape[-1]
    ws = w * he_scale(w.shape[1] * k * k, gain)
    if min(x.shape[2:]) >= 128:
        wp = F.pad(ws, [1, 1, 1, 1])
        wp = (wp[:, :, 1:, 1:] + wp[:, :, :-1, 1:] + wp[:, :, 1:, :-1] + wp[:, :, :-1, :-1]) * 0.25
        y = F.conv2d(x, wp, stride=2, padding=1)
    else:
        y = avg_pool2(F.conv2d(x, ws, None, padding=1))
    return y + b.view(1, -1, 1, 1)


def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.InstanceNorm2d defaults: biased variance over HxW, eps 1e-5, no affine.  CustomLayers.py:233."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps)


def layer_epilogue(x: Tensor, noise_w: Tensor, noise: Optional[Tensor], style_w: Tensor, style_b: Tensor,
                   dlatent: Tensor) -> Tensor:
    """noise -> LeakyReLU -> InstanceNorm -> StyleMod.  models/CustomLayers.py:191-248.

    ``noise`` is [B,1,H,W]; None means "all zeros" (reference would draw randn; tests always pass it).
    """
    if noise is not None:
        x = x + noise_w.view(1, -1, 1, 1) * noise
    x = instance_norm(lrelu(x))
    style = eq_linear(dlatent, style_w, style_b, gain=1.0)            # CustomLayers.py:206-211
    c = x.shape[1]
    s = style.view(-1, 2, c, 1, 1)
    return x * (s[:, 0] + 1.0) + s[:, 1]                              # CustomLayers.py:213-216


def minibatch_stddev(x: Tensor, group_size: int = 4) -> Tensor:
    """models/CustomLayers.py:294-305 with num_new_features=1."""
    b, c, h, w = x.shape
    g = min(group_size, b)
    y = x.reshape(g, -1, 1, c, h, w)
    y = y - y.mean(0, keepdim=True)
    y = (y * y).mean(0, keepdim=True)
    y = torch.sqrt(y + 1e-8)
    y = y.mean(dim=(3, 4, 5), keepdim=True).squeeze(3)
    y = y.expand(g, -1, -1, h, w).clone().reshape(b, 1, h, w)
    return torch.cat([x, y], dim=1)


# --------------------------------------------------------------------------- #
# networks
# --------------------------------------------------------------------------- #
SQRT2 = math.sqrt(2.0)


def g_mapping(sd: SD, z: Tensor, arch: Arch, prefix: str = "g_mapping.map.") -> Tensor:
    """PixelNorm -> mapping_layers x (dense lrmul=.01 + lrelu).  models/GAN.py:73-100.  Returns [B, dlatent]."""
    x = pixel_norm(z)
    for i in range(arch.mapping_layers):
        x = lrelu(eq_linear(x, sd[f"{prefix}dense{i}.weight"], sd[f"{prefix}dense{i}.bias"], SQRT2, lrmul=0.01))
    return x


def noise_shapes(arch: Arch, batch: int, depth: int) -> List[Tuple[int, int, int, int]]:
    """Shapes of the noise tensors consumed, in call order, by g_synthesis(depth)."""
    shapes = [(batch, 1, 4, 4)] * 2
    for i in range(depth):
        r = 2 ** (i + 3)
        shapes += [(batch, 1, r, r)] * 2
    return shapes


def g_synthesis(sd: SD, dlatents: Tensor, depth: int, alpha: float, arch: Arch,
                noises: Optional[Sequence[Tensor]] = None, prefix: str = "g_synthesis.") -> Tensor:
    """models/GAN.py:175-208 ('linear' and 'fixed'), Blocks.py:47-60, 83-88.  dlatents: [B, L, D]."""
    assert depth < arch.depth, "Requested output depth cannot be produced"
    nz = iter(noises) if noises is not None else None

    def nxt():
        return next(nz) if nz is not None else None

    def epi(x, p, dl):
        return layer_epilogue(x, sd[p + "top_epi.noise.weight"], nxt(), sd[p + "style_mod.lin.weight"],
                              sd[p + "style_mod.lin.bias"], dl)

    b = dlatents.shape[0]
    p = prefix + "init_block."
    x = sd[p + "const"].expand(b, -1, -1, -1) + sd[p + "bias"].view(1, -1, 1, 1)       # Blocks.py:51-52
    x = epi(x, p + "epi1.", dlatents[:, 0])
    x = eq_conv(x, sd[p + "conv.weight"], sd[p + "conv.bias"], SQRT2)
    x = epi(x, p + "epi2.", dlatents[:, 1])

    def block(x, i):
        q = f"{prefix}blocks.{i}."
        x = up_conv(x, sd[q + "conv0_up.weight"], sd[q + "conv0_up.bias"], SQRT2, arch.blur)
        x = epi(x, q + "epi1.", dlatents[:, 2 * (i + 1)])
        x = eq_conv(x, sd[q + "conv1.weight"], sd[q + "conv1.bias"], SQRT2)
        return epi(x, q + "epi2.", dlatents[:, 2 * (i + 1) + 1])

    def to_rgb(x, j):
        return eq_conv(x, sd[f"{prefix}to_rgb.{j}.weight"], sd[f"{prefix}to_rgb.{j}.bias"], 1.0)

    if arch.structure == "fixed":
        for i in range(arch.depth - 1):
            x = block(x, i)
        return to_rgb(x, arch.depth - 1)
    if depth == 0:
        return to_rgb(x, 0)
    for i in range(depth - 1):
        x = block(x, i)
    residual = to_rgb(nearest_up2(x), depth - 1)                                         # GAN.py:199
    straight = to_rgb(block(x, depth - 1), depth)                                        # GAN.py:200
    return alpha * straight + (1 - alpha) * residual                                     # GAN.py:202


def truncate(dlatents: Tensor, avg: Tensor, psi: float, cutoff: int) -> Tensor:
    """models/CustomLayers.py:319-323."""
    interp = torch.lerp(avg, dlatents, psi)
    mask = (torch.arange(dlatents.shape[1], device=dlatents.device) < cutoff).view(1, -1, 1)
    return torch.where(mask, interp, dlatents)


def generator_forward(sd: SD, z: Tensor, depth: int, alpha: float, arch: Arch, *, training: bool = True,
                      z2: Optional[Tensor] = None, mixing_cutoff: Optional[int] = None,
                      noises: Optional[Sequence[Tensor]] = None, update_avg: bool = True) -> Tensor:
    """models/GAN.py:254-297.  In training mode: W-average update from sample 0, style mixing with the given
    second latent ``z2`` and layer cutoff, then truncation.  ``sd['truncation.avg_latent']`` is updated in place
    (it is a buffer in the reference).  The host-side random draws (z2: GAN.py:282, cutoff: GAN.py:287-288) are
    inputs here so the oracle is a pure function; `draw_mixing` reproduces them."""
    w = g_mapping(sd, z, arch)
    dl = w.unsqueeze(1).expand(-1, arch.num_layers, -1)
    has_trunc = arch.truncation_psi > 0
    if training:
        if has_trunc and update_avg:
            avg = sd["truncation.avg_latent"]
            avg.copy_(arch.dlatent_avg_beta * avg + (1.0 - arch.dlatent_avg_beta) * w[0].detach())   # GAN.py:278
        if arch.style_mixing_prob is not None and arch.style_mixing_prob > 0:
            assert z2 is not None and mixing_cutoff is not None
            w2 = g_mapping(sd, z2, arch)
            dl2 = w2.unsqueeze(1).expand(-1, arch.num_layers, -1)
            layer_idx = torch.arange(arch.num_layers, device=dl.device).view(1, -1, 1)
            dl = torch.where(layer_idx < mixing_cutoff, dl, dl2)                                       # GAN.py:289
        if has_trunc:
            dl = truncate(dl, sd["truncation.avg_latent"], arch.truncation_psi, arch.truncation_cutoff)
    return g_synthesis(sd, dl, depth, alpha, arch, noises)


def draw_mixing(arch: Arch, z_shape, depth: int) -> Tuple[Tensor, int]:
    """Reproduce the reference's host RNG consumption in one training-mode Generator.forward:
    torch.randn(z.shape) on the CPU generator (GAN.py:282), then random.random()/random.randint (GAN.py:287-288)."""
    z2 = torch.randn(z_shape)
    cur_layers = 2 * (depth + 1)
    if _pyrandom.random() < arch.style_mixing_prob:
        cutoff = _pyrandom.randint(1, cur_layers)
    else:
        cutoff = cur_layers
    return z2, cutoff


def discriminator(sd: SD, img: Tensor, depth: int, alpha: float, arch: Arch) -> Tensor:
    """models/GAN.py:384-444 (unconditional), Blocks.py:91-146.  Returns [B,1]."""
    assert depth < arch.depth, "Requested output depth cannot be produced"

    def from_rgb(x, j):
        return eq_conv(x, sd[f"from_rgb.{j}.weight"], sd[f"from_rgb.{j}.bias"], SQRT2)

    def block(x, i):
        q = f"blocks.{i}."
        x = lrelu(eq_conv(x, sd[q + "conv0.weight"], sd[q + "conv0.bias"], SQRT2))
        if arch.blur:
            x = blur3(x)
        return lrelu(down_conv(x, sd[q + "conv1_down.weight"], sd[q + "conv1_down.bias"], SQRT2))

    D = arch.depth
    if arch.structure == "fixed":
        x = from_rgb(img, 0)
        for i in range(D - 1):
            x = block(x, i)
    elif depth > 0:
        residual = from_rgb(avg_pool2(img), D - depth)                                    # GAN.py:423-424
        straight = block(from_rgb(img, D - depth - 1), D - depth - 1)                     # GAN.py:425-426
        x = alpha * straight + (1 - alpha) * residual
        for i in range(D - depth, D - 1):
            x = block(x, i)
    else:
        x = from_rgb(img, D - 1)

    if arch.mbstd_group_size > 1:
        x = minibatch_stddev(x, arch.mbstd_group_size)
    x = lrelu(eq_conv(x, sd["final_block.conv.weight"], sd["final_block.conv.bias"], SQRT2))
    x = x.reshape(x.shape[0], -1)
    x = lrelu(eq_linear(x, sd["final_block.dense0.weight"], sd["final_block.dense0.bias"], SQRT2))
    return eq_linear(x, sd["final_block.dense1.weight"], sd["final_block.dense1.bias"], 1.0)


# --------------------------------------------------------------------------- #
# training step
# --------------------------------------------------------------------------- #
def progressive_down_sampling(real: Tensor, arch: Arch, depth: int, alpha: float) -> Tensor:
    """models/GAN.py:557-589."""
    if arch.structure == "fixed":
        return real
    f = int(2 ** (arch.depth - depth - 1))
    pf = int(2 ** (arch.depth - depth))
    ds = F.avg_pool2d(real, f) if f > 1 else real
    prior = nearest_up2(F.avg_pool2d(real, pf)) if depth > 0 else ds
    return alpha * ds + (1 - alpha) * prior


def softplus(x: Tensor) -> Tensor:
    return F.softplus(x)


def r1_penalty(sd_d: SD, real: Tensor, depth: int, alpha: float, arch: Arch) -> Tensor:
    """models/Losses.py:197-210: sum over batch and pixels of (d D(real)/d real)^2, graph kept."""
    real = real.detach().requires_grad_(True)
    logit = discriminator(sd_d, real, depth, alpha, arch)
    (g,) = torch.autograd.grad(logit, real, grad_outputs=torch.ones_like(logit), create_graph=True)
    return (g * g).sum()


def d_loss_logistic(sd_d: SD, real: Tensor, fake: Tensor, depth: int, alpha: float, arch: Arch,
                    r1_gamma: float = 10.0) -> Tensor:
    """models/Losses.py:213-224."""
    r = discriminator(sd_d, real, depth, alpha, arch)
    f = discriminator(sd_d, fake, depth, alpha, arch)
    loss = softplus(f).mean() + softplus(-r).mean()
    if r1_gamma != 0.0:
        loss = loss + r1_penalty(sd_d, real, depth, alpha, arch) * (r1_gamma * 0.5)
    return loss


def g_loss_logistic(sd_d: SD, fake: Tensor, depth: int, alpha: float, arch: Arch) -> Tensor:
    """models/Losses.py:226-229."""
    return softplus(-discriminator(sd_d, fake, depth, alpha, arch)).mean()


def d_loss_hinge(sd_d, real, fake, depth, alpha, arch):
    """models/Losses.py:141-148."""
    r = discriminator(sd_d, real, depth, alpha, arch)
    f = discriminator(sd_d, fake, depth, alpha, arch)
    return F.relu(1 - r).mean() + F.relu(1 + f).mean()


def g_loss_hinge(sd_d, fake, depth, alpha, arch):
    """models/Losses.py:150-151."""
    return -discriminator(sd_d, fake, depth, alpha, arch).mean()


def d_loss_rahinge(sd_d, real, fake, depth, alpha, arch):
    """models/Losses.py:159-173."""
    r = discriminator(sd_d, real, depth, alpha, arch)
    f = discriminator(sd_d, fake, depth, alpha, arch)
    return F.relu(1 - (r - f.mean())).mean() + F.relu(1 + (f - r.mean())).mean()


def g_loss_rahinge(sd_d, real, fake, depth, alpha, arch):
    """models/Losses.py:175-189."""
    r = discriminator(sd_d, real, depth, alpha, arch)
    f = discriminator(sd_d, fake, depth, alpha, arch)
    return F.relu(1 + (r - f.mean())).mean() + F.relu(1 - (f - r.mean())).mean()


class AdamState:
    """torch.optim.Adam (no amsgrad, no weight decay) restated; models/GAN.py:529-533, config.py:79-92.

    As in torch, a parameter whose gradient is None (a block unused at the current depth) is skipped and keeps
    its own step counter; state is per parameter."""

    def __init__(self, names: Sequence[str], params: SD, lr=0.003, beta1=0.0, beta2=0.99, eps=1e-8):
        self.names = list(names)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.steps = {k: 0 for k in self.names}
        self.m = {k: torch.zeros_like(params[k]) for k in self.names}
        self.v = {k: torch.zeros_like(params[k]) for k in self.names}

    @torch.no_grad()
    def update(self, params: SD, grads: Dict[str, Optional[Tensor]]):
        for k in self.names:
            g = grads.get(k)
            if g is None:
                continue
            self.steps[k] += 1
            t = self.steps[k]
            bc1 = 1.0 - self.b1 ** t
            bc2 = 1.0 - self.b2 ** t
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            params[k].addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def clip_grad_norm(grads: Dict[str, Tensor], names: Sequence[str], max_norm: float = 10.0) -> float:
    """nn.utils.clip_grad_norm_ (L2, eps 1e-6, coefficient clamped to 1).  models/GAN.py:651."""
    total = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in names)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for k in names:
        grads[k].mul_(coef)
    return float(total)


def param_names(sd: SD) -> List[str]:
    """Parameters (not buffers) in registration order: buffers are the blur kernels and avg_latent."""
    return [k for k in sd if not (k.endswith(".kernel") or k == "truncation.avg_latent")]


class TrainState:
    """Everything optimize_discriminator/optimize_generator mutate.  models/GAN.py:447-659."""

    def __init__(self, arch: Arch, g_sd: SD, d_sd: SD, use_ema: bool = True, ema_decay: float = 0.999,
                 lr=0.003, beta1=0.0, beta2=0.99, eps=1e-8):
        self.arch = arch
        self.g = {k: v.detach().clone() for k, v in g_sd.items()}
        self.d = {k: v.detach().clone() for k, v in d_sd.items()}
        self.g_names = param_names(self.g)
        self.d_names = param_names(self.d)
        self.g_opt = AdamState(self.g_names, self.g, lr, beta1, beta2, eps)
        self.d_opt = AdamState(self.d_names, self.d, lr, beta1, beta2, eps)
        self.use_ema = use_ema
        self.ema_decay = ema_decay
        self.shadow = {k: v.clone() for k, v in self.g.items()} if use_ema else None   # GAN.py:521-527

    def _leaf(self, sd: SD, names: Sequence[str]) -> SD:
        out = dict(sd)
        for k in names:
            out[k] = sd[k].detach().requires_grad_(True)
        return out


def train_step(st: TrainState, real_batch: Tensor, z: Tensor, depth: int, alpha: float,
               mix_d: Tuple[Tensor, int], mix_g: Tuple[Tensor, int],
               noises_d: Optional[Sequence[Tensor]], noises_g: Optional[Sequence[Tensor]],
               loss: str = "logistic") -> Tuple[float, float]:
    """One optimize_discriminator + optimize_generator (d_repeats=1).  models/GAN.py:591-659.

    mix_* = (z2, cutoff) for the style-mixing branch of each of the two Generator.forward calls;
    noises_* = per-layer noise tensors for each call.
    """
    arch = st.arch
    real = progressive_down_sampling(real_batch, arch, depth, alpha)                     # GAN.py:602

    # ---- discriminator step (GAN.py:605-620) ----
    with torch.no_grad():
        fake = generator_forward(st.g, z, depth, alpha, arch, training=True, z2=mix_d[0],
                                 mixing_cutoff=mix_d[1], noises=noises_d)
    d_leaf = st._leaf(st.d, st.d_names)
    if loss == "logistic":
        d_loss = d_loss_logistic(d_leaf, real, fake, depth, alpha, arch)
    elif loss == "hinge":
        d_loss = d_loss_hinge(d_leaf, real, fake, depth, alpha, arch)
    elif loss == "relativistic-hinge":
        d_loss = d_loss_rahinge(d_leaf, real, fake, depth, alpha, arch)
    else:
        raise AssertionError("Unknown loss function")
    grads = torch.autograd.grad(d_loss, [d_leaf[k] for k in st.d_names], allow_unused=True)
    st.d_opt.update(st.d, dict(zip(st.d_names, grads)))

    # ---- generator step (GAN.py:635-659) ----
    g_leaf = st._leaf(st.g, st.g_names)
    fake = generator_forward(g_leaf, z, depth, alpha, arch, training=True, z2=mix_g[0],
                             mixing_cutoff=mix_g[1], noises=noises_g)
    if loss == "logistic":
        g_loss = g_loss_logistic(st.d, fake, depth, alpha, arch)
    elif loss == "hinge":
        g_loss = g_loss_hinge(st.d, fake, depth, alpha, arch)
    else:
        g_loss = g_loss_rahinge(st.d, real, fake, depth, alpha, arch)
    grads = torch.autograd.grad(g_loss, [g_leaf[k] for k in st.g_names], allow_unused=True)
    used = [k for k, g in zip(st.g_names, grads) if g is not None]
    gd_used = {k: g.clone() for k, g in zip(st.g_names, grads) if g is not None}
    clip_grad_norm(gd_used, used, 10.0)                                                   # GAN.py:651
    st.g_opt.update(st.g, gd_used)
    if st.use_ema:                                                                        # models/__init__.py:33-36
        with torch.no_grad():
            for k in st.g_names:
                st.shadow[k].copy_(st.ema_decay * st.shadow[k] + (1.0 - st.ema_decay) * st.g[k])
    return float(d_loss.detach()), float(g_loss.detach())


# --------------------------------------------------------------------------- #
# random initialisation with the reference's distributions (for benchmarks / fixtures without the reference)
# --------------------------------------------------------------------------- #
def init_generator(arch: Arch, gen: Optional[torch.Generator] = None) -> SD:
    """Keys/shapes per SURVEY Appendix B; distributions per CustomLayers.py:84-95,120-132,188, Blocks.py:33-35."""
    def rn(*s):
        return torch.randn(*s, generator=gen)
    sd: SD = {}
    L, Dl, Fm = arch.latent_size, arch.dlatent_size, arch.mapping_fmaps
    for i in range(arch.mapping_layers):
        fin = L if i == 0 else Fm
        fout = Dl if i == arch.mapping_layers - 1 else Fm
        sd[f"g_mapping.map.dense{i}.weight"] = rn(fout, fin) / 0.01
        sd[f"g_mapping.map.dense{i}.bias"] = torch.zeros(fout)
    c0 = arch.nf(1)
    p = "g_synthesis.init_block."
    sd[p + "const"] = torch.ones(1, c0, 4, 4)
    sd[p + "bias"] = torch.ones(c0)

    def epi(q, c):
        sd[q + "top_epi.noise.weight"] = torch.zeros(c)
        sd[q + "style_mod.lin.weight"] = rn(2 * c, Dl)
        sd[q + "style_mod.lin.bias"] = torch.zeros(2 * c)
    epi(p + "epi1.", c0)
    sd[p + "conv.weight"] = rn(c0, c0, 3, 3)
    sd[p + "conv.bias"] = torch.zeros(c0)
    epi(p + "epi2.", c0)
    k1 = torch.tensor([1.0, 2.0, 1.0])
    blur_k = ((k1[:, None] * k1[None, :]) / 16.0)[None, None]
    for i in range(arch.depth - 1):
        ci, co = arch.g_block_channels(i)
        q = f"g_synthesis.blocks.{i}."
        sd[q + "conv0_up.weight"] = rn(co, ci, 3, 3)
        sd[q + "conv0_up.bias"] = torch.zeros(co)
        if arch.blur:
            sd[q + "conv0_up.intermediate.kernel"] = blur_k.clone()
        epi(q + "epi1.", co)
        sd[q + "conv1.weight"] = rn(co, co, 3, 3)
        sd[q + "conv1.bias"] = torch.zeros(co)
        epi(q + "epi2.", co)
    for j in range(arch.depth):
        c = arch.nf(j + 1)
        sd[f"g_synthesis.to_rgb.{j}.weight"] = rn(arch.num_channels, c, 1, 1)
        sd[f"g_synthesis.to_rgb.{j}.bias"] = torch.zeros(arch.num_channels)
    if arch.truncation_psi > 0:
        sd["truncation.avg_latent"] = torch.zeros(Dl)
    return sd


def init_discriminator(arch: Arch, gen: Optional[torch.Generator] = None) -> SD:
    def rn(*s):
        return torch.randn(*s, generator=gen)
    sd: SD = {}
    k1 = torch.tensor([1.0, 2.0, 1.0])
    blur_k = ((k1[:, None] * k1[None, :]) / 16.0)[None, None]
    for i in range(arch.depth - 1):
        ci, co = arch.d_block_channels(i)
        q = f"blocks.{i}."
        sd[q + "conv0.weight"] = rn(ci, ci, 3, 3)
        sd[q + "conv0.bias"] = torch.zeros(ci)
        if arch.blur:
            sd[q + "blur.kernel"] = blur_k.clone()
        sd[q + "conv1_down.weight"] = rn(co, ci, 3, 3)
        sd[q + "conv1_down.bias"] = torch.zeros(co)
        sd[q + "conv1_down.downscale.blur.kernel"] = torch.full((1, 1, 2, 2), 0.25)
    c = arch.nf(2)
    sd["final_block.conv.weight"] = rn(c, c + 1, 3, 3)
    sd["final_block.conv.bias"] = torch.zeros(c)
    sd["final_block.dense0.weight"] = rn(c, c * 16)
    sd["final_block.dense0.bias"] = torch.zeros(c)
    sd["final_block.dense1.weight"] = rn(1, c)
    sd["final_block.dense1.bias"] = torch.zeros(1)
    for j in range(arch.depth):
        res = arch.res_log2 - j
        cj = arch.nf(res - 1)
        sd[f"from_rgb.{j}.weight"] = rn(cj, arch.num_channels, 1, 1)
        sd[f"from_rgb.{j}.bias"] = torch.zeros(cj)
    return sd


# --------------------------------------------------------------------------- #
# seeded full-width cases (fixtures too large to commit are re-derived from a seed)
# --------------------------------------------------------------------------- #
def seeded_case(arch: Arch, batch: int, depth: int, seed: int) -> dict:
    """Every INPUT of a full-width forward/backward fixture, derived from one CPU generator seed: initial state
    dicts (reference distributions; noise weights ~N(0,1) and biases ~0.1 N(0,1) so those paths carry signal,
    SURVEY 8(d)), per-layer noise, latents, second latents + mixing cut-off, real images.  oracle/make_golden.py
    feeds exactly this to the live reference and commits only the OUTPUTS; the tests rebuild the inputs on the GPU box
    with the same call (torch's CPU Philox/MT stream is a function of the seed and the torch build, which is the
    same image on both sides)."""
    gen = torch.Generator().manual_seed(seed)
    g_sd, d_sd = init_generator(arch, gen), init_discriminator(arch, gen)
    for sd in (g_sd, d_sd):
        for k in sd:
            if k.endswith("noise.weight"):
                sd[k] = torch.randn(sd[k].shape, generator=gen)
            elif k.endswith(".bias") and "init_block.bias" not in k:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
    noises = [torch.randn(s, generator=gen) for s in noise_shapes(arch, batch, depth)]
    z = torch.randn(batch, arch.latent_size, generator=gen)
    z2 = torch.randn(batch, arch.latent_size, generator=gen)
    cutoff = int(torch.randint(1, 2 * (depth + 1) + 1, (1,), generator=gen))
    res = 2 ** (depth + 2)
    real = torch.randn(batch, arch.num_channels, res, res, generator=gen)
    return dict(g_sd=g_sd, d_sd=d_sd, noises=noises, z=z, z2=z2, cutoff=cutoff, real=real)
