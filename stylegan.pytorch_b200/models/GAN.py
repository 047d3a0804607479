"""StyleGAN networks and the training wrapper with the reference's public API
(huangzh13/StyleGAN.pytorch models/GAN.py): ``GMapping``, ``GSynthesis``, ``Generator``, ``Discriminator``,
``StyleGAN`` -- same constructor signatures, attribute names, state_dict keys, parameter registration order and
checkpoint files, so the reference's train.py / generate_*.py run against this package unchanged.

The per-step hot path (optimize_discriminator / optimize_generator) runs on the sgb200 CUDA kernels and, after two
eager calls, as a replayed CUDA graph (models/_graphs.py); under torch.distributed (one process per GPU, NCCL) the
gradients of each optimizer are all-reduced in one flat bucket after its backward (models/_ddp.py) and consumed in
place by the fused Adam step."""
import copy
import datetime
import os
import random
import time
import timeit
import warnings
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import Losses, _host_facade, _ops, update_average
from ._optim import FusedAdam
from .Blocks import DiscriminatorBlock, DiscriminatorTop, GSynthesisBlock, InputBlock
from .CustomLayers import EqualizedConv2d, EqualizedLinear, PixelNormLayer, Truncation, to_nchw, to_nhwc


def _activation(name):
    if name == "lrelu":
        return nn.LeakyReLU(negative_slope=0.2), np.sqrt(2)
    if name == "relu":
        raise NotImplementedError("sgb200: only the 'lrelu' nonlinearity (the StyleGAN default) has kernels")
    raise KeyError(name)


def _nf(stage, fmap_base, fmap_decay, fmap_max):
    return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)


class GMapping(nn.Module):
    """Z -> W mapping MLP (reference GAN.py:37-100)."""

    def __init__(self, latent_size=512, dlatent_size=512, dlatent_broadcast=None, mapping_layers=8, mapping_fmaps=512,
                 mapping_lrmul=0.01, mapping_nonlinearity='lrelu', use_wscale=True, normalize_latents=True, **kwargs):
        super().__init__()
        self.latent_size = latent_size
        self.mapping_fmaps = mapping_fmaps
        self.dlatent_size = dlatent_size
        self.dlatent_broadcast = dlatent_broadcast
        act, gain = _activation(mapping_nonlinearity)
        layers = []
        if normalize_latents:
            layers.append(('pixel_norm', PixelNormLayer()))
        for idx in range(mapping_layers):
            fin = self.latent_size if idx == 0 else self.mapping_fmaps
            fout = self.dlatent_size if idx == mapping_layers - 1 and idx > 0 else self.mapping_fmaps
            layers.append(('dense{:d}'.format(idx),
                           EqualizedLinear(fin, fout, gain=gain, lrmul=mapping_lrmul, use_wscale=use_wscale)))
            layers.append(('dense{:d}_act'.format(idx), act))
        self.map = nn.Sequential(OrderedDict(layers))

    def forward(self, x):
        if _host_facade.wants_facade(self, x):          # CPU module + CPU latents (generate_* scripts)
            return _host_facade.call(self, x)
        _ops._require_cuda(x)
        x = x.float()
        for name, m in self.map.named_children():
            if isinstance(m, EqualizedLinear):
                x = m(x, act=True)              # dense + LeakyReLU fused
            elif isinstance(m, PixelNormLayer):
                x = m(x)
        if self.dlatent_broadcast is not None:
            x = x.unsqueeze(1).expand(-1, self.dlatent_broadcast, -1)
        return x


class GSynthesis(nn.Module):
    """Progressive synthesis network (reference GAN.py:103-208)."""

    def __init__(self, dlatent_size=512, num_channels=3, resolution=1024, fmap_base=8192, fmap_decay=1.0, fmap_max=512,
                 use_styles=True, const_input_layer=True, use_noise=True, nonlinearity='lrelu', use_wscale=True,
                 use_pixel_norm=False, use_instance_norm=True, blur_filter=None, structure='linear', **kwargs):
        super().__init__()

        def nf(stage):
            return _nf(stage, fmap_base, fmap_decay, fmap_max)

        self.structure = structure
        resolution_log2 = int(np.log2(resolution))
        assert resolution == 2 ** resolution_log2 and resolution >= 4
        self.depth = resolution_log2 - 1
        self.num_layers = resolution_log2 * 2 - 2
        self.num_styles = self.num_layers if use_styles else 1
        act, gain = _activation(nonlinearity)

        self.init_block = InputBlock(nf(1), dlatent_size, const_input_layer, gain, use_wscale, use_noise,
                                     use_pixel_norm, use_instance_norm, use_styles, act)
        rgb = [EqualizedConv2d(nf(1), num_channels, 1, gain=1, use_wscale=use_wscale)]
        blocks = []
        for res in range(3, resolution_log2 + 1):
            blocks.append(GSynthesisBlock(nf(res - 2), nf(res - 1), blur_filter, dlatent_size, gain, use_wscale,
                                          use_noise, use_pixel_norm, use_instance_norm, use_styles, act))
            rgb.append(EqualizedConv2d(nf(res - 1), num_channels, 1, gain=1, use_wscale=use_wscale))
        self.blocks = nn.ModuleList(blocks)
        self.to_rgb = nn.ModuleList(rgb)
        self.temporaryUpsampler = lambda x: to_nchw(_upsample_feature(to_nhwc(x)))

    def _rgb(self, x_hi, j_hi, a_hi=1.0, x_lo=None, j_lo=None, a_lo=0.0):
        hi = self.to_rgb[j_hi]
        if x_lo is None:
            return _ops.ToRgb.apply(x_hi, hi.weight, hi.bias, hi.w_mul, a_hi, None, None, None, 1.0, 0.0)
        lo = self.to_rgb[j_lo]
        return _ops.ToRgb.apply(x_hi, hi.weight, hi.bias, hi.w_mul, a_hi, x_lo, lo.weight, lo.bias, lo.w_mul, a_lo)

    def _precompute_styles(self, dl, n_blocks):
        """All StyleMod affines this pass will need, in one grouped launch, handed to their LayerEpilogues."""
        epis = [(0, self.init_block.epi1), (1, self.init_block.epi2)]
        for i in range(n_blocks):
            epis += [(2 * (i + 1), self.blocks[i].epi1), (2 * (i + 1) + 1, self.blocks[i].epi2)]
        epis = [(l, e) for l, e in epis if e.style_mod is not None and e._fusable]
        if not epis or dl.dim() != 3 or not _ops.StyleAffines.supported(dl):
            return
        if not dl.is_contiguous():
            dl = dl.contiguous()
        params, cfg = [], []
        for _, e in epis:
            lin = e.style_mod.lin
            params += [lin.weight, lin.bias]
            cfg.append((lin.w_mul, lin.b_mul if lin.bias is not None else 1.0))
        styles = _ops.StyleAffines.apply(dl, tuple(l for l, _ in epis), tuple(cfg), *params)
        for (_, e), s in zip(epis, styles):
            e._style_pre = s

    def forward(self, dlatents_in, depth=0, alpha=0., labels_in=None):
        """dlatents_in [B, num_layers, dlatent] -> images [B, C, R, R] float32 (NCHW)."""
        assert depth < self.depth, "Requested output depth cannot be produced"
        if _host_facade.wants_facade(self, dlatents_in):
            return _host_facade.call(self, dlatents_in, depth=depth, alpha=alpha, labels_in=labels_in)
        _ops._require_cuda(dlatents_in)
        dl = dlatents_in.float()
        self._precompute_styles(dl, len(self.blocks) if self.structure == 'fixed' else depth)
        if self.structure == 'fixed':
            x = self.init_block.forward_nhwc(dl[:, 0:2])
            for i, block in enumerate(self.blocks):
                x = block.forward_nhwc(x, dl[:, 2 * (i + 1):2 * (i + 2)])
            return self._rgb(x, len(self.to_rgb) - 1)
        if self.structure == 'linear':
            x = self.init_block.forward_nhwc(dl[:, 0:2])
            if depth == 0:
                return self._rgb(x, 0)
            for i in range(depth - 1):
                x = self.blocks[i].forward_nhwc(x, dl[:, 2 * (i + 1):2 * (i + 2)])
            x_hi = self.blocks[depth - 1].forward_nhwc(x, dl[:, 2 * depth:2 * (depth + 1)])
            # straight and residual branches + fade-in (reference GAN.py:199-202).  Both branches with unit weights -- the residual one at the low resolution: the 1x1 convolution commutes with
            # nearest up-sampling (reference GAN.py:199 up-samples the features first) -- then one blend whose coefficient
            # pair is either two floats or, under CUDA-graph replay, read from device memory by the same kernel.  Eager
            # and replayed steps therefore execute identical arithmetic.
            hi = self._rgb(x_hi, depth, 1.0)
            lo = _ops.UpsampleNCHW.apply(self._rgb(x, depth - 1, 1.0), 2, 1.0)
            return _ops.Axpby.apply(hi, lo, *_ops.lerp_coefs(alpha))
        raise KeyError("Unknown structure: ", self.structure)


def _upsample_feature(x_nhwc):
    n, h, w, c = x_nhwc.shape
    planes = _ops.Cast.apply(x_nhwc, torch.float32).permute(0, 3, 1, 2).contiguous()
    up = _ops.UpsampleNCHW.apply(planes, 2, 1.0)
    return _ops.Cast.apply(up.permute(0, 2, 3, 1).contiguous(), x_nhwc.dtype)


def draw_style_mixing(shape, depth, prob):
    """The host RNG draws of one training-mode Generator.forward, in the reference's order (GAN.py:282, 287-288):
    torch.randn on the CPU generator, then random.random() / random.randint().  -> (latents2 on the host, cut-off)"""
    latents2 = torch.randn(shape)
    cur_layers = 2 * (depth + 1)
    cutoff = random.randint(1, cur_layers) if random.random() < prob else cur_layers
    return latents2, cutoff


class Generator(nn.Module):
    """Mapping + synthesis + truncation / style mixing (reference GAN.py:211-297)."""

    def __init__(self, resolution, latent_size=512, dlatent_size=512, conditional=False, n_classes=0,
                 truncation_psi=0.7, truncation_cutoff=8, dlatent_avg_beta=0.995, style_mixing_prob=0.9, **kwargs):
        super(Generator, self).__init__()
        if conditional:
            assert n_classes > 0, "Conditional generation requires n_class > 0"
            self.class_embedding = nn.Embedding(n_classes, latent_size)
            latent_size *= 2
        self.conditional = conditional
        self.style_mixing_prob = style_mixing_prob
        self.num_layers = (int(np.log2(resolution)) - 1) * 2
        self.g_mapping = GMapping(latent_size, dlatent_size, dlatent_broadcast=self.num_layers, **kwargs)
        self.g_synthesis = GSynthesis(resolution=resolution, **kwargs)
        if truncation_psi > 0:
            self.truncation = Truncation(avg_latent=torch.zeros(dlatent_size), max_layer=truncation_cutoff,
                                         threshold=truncation_psi, beta=dlatent_avg_beta)
        else:
            self.truncation = None
        self._avg_sync = None           # set by StyleGAN under data parallelism (models/_ddp.py)

    def forward(self, latents_in, depth, alpha, labels_in=None):
        if _host_facade.wants_facade(self, latents_in, labels_in):
            return _host_facade.call(self, latents_in, depth, alpha, labels_in=labels_in)
        if not self.conditional:
            if labels_in is not None:
                warnings.warn("Generator is unconditional, labels_in will be ignored")
        else:
            assert labels_in is not None, "Conditional discriminatin requires labels"
            latents_in = torch.cat([latents_in, self.class_embedding(labels_in)], 1)

        w1 = self.g_mapping(latents_in)
        w1 = w1[:, 0] if w1.dim() == 3 else w1
        L = self.num_layers
        if not self.training:
            dlatents = w1.unsqueeze(1).expand(-1, L, -1)
            return self.g_synthesis(dlatents, depth, alpha)

        # training mode (reference GAN.py:274-293): W-average update from sample 0, style mixing, truncation.
        if self.truncation is not None:
            self.truncation.update(w1[0])
            if self._avg_sync is not None:
                # data parallel: every rank folded its own sample 0 into its copy (reference GAN.py:278 knows one
                # process only); the replicas' buffers are averaged so that they stay identical (the EMA is linear,
                # so this equals one EMA over the mean of the ranks' samples)
                self._avg_sync(self.truncation.avg_latent)
        w2, cutoff = None, L
        if self.style_mixing_prob is not None and self.style_mixing_prob > 0:
            # host RNG consumption order identical to the reference: torch.randn on the CPU generator, then
            # random.random() / random.randint()
            cap = _ops.capture_ctx()
            if cap is not None:
                # CUDA-graph capture: the host draws are made by the replay driver (same calls, same order) and arrive
                # through device buffers
                latents2, cutoff = cap.next_mixing(latents_in.shape)
            else:
                latents2, cutoff = draw_style_mixing(latents_in.shape, depth, self.style_mixing_prob)
                latents2 = latents2.to(latents_in.device)
            w2 = self.g_mapping(latents2)
            w2 = w2[:, 0] if w2.dim() == 3 else w2
        if self.truncation is not None:
            dlatents = _ops.Dlatents.apply(w1, w2, self.truncation.avg_latent, L, cutoff, self.truncation.threshold,
                                           self.truncation.max_layer)
        else:
            dlatents = _ops.Dlatents.apply(w1, w2, None, L, cutoff, 1.0, 0)
        return self.g_synthesis(dlatents, depth, alpha)


class Discriminator(nn.Module):
    """Progressive discriminator (reference GAN.py:300-444)."""

    def __init__(self, resolution, num_channels=3, conditional=False, n_classes=0, fmap_base=8192, fmap_decay=1.0,
                 fmap_max=512, nonlinearity='lrelu', use_wscale=True, mbstd_group_size=4, mbstd_num_features=1,
                 blur_filter=None, structure='linear', **kwargs):
        super(Discriminator, self).__init__()
        if conditional:
            assert n_classes > 0, "Conditional Discriminator requires n_class > 0"
            num_channels *= 2
            self.embeddings = []

        def nf(stage):
            return _nf(stage, fmap_base, fmap_decay, fmap_max)

        self.conditional = conditional
        self.mbstd_num_features = mbstd_num_features
        self.mbstd_group_size = mbstd_group_size
        self.structure = structure
        resolution_log2 = int(np.log2(resolution))
        assert resolution == 2 ** resolution_log2 and resolution >= 4
        self.depth = resolution_log2 - 1
        act, gain = _activation(nonlinearity)

        blocks, from_rgb = [], []
        for res in range(resolution_log2, 2, -1):
            blocks.append(DiscriminatorBlock(nf(res - 1), nf(res - 2), gain=gain, use_wscale=use_wscale,
                                             activation_layer=act, blur_kernel=blur_filter))
            from_rgb.append(EqualizedConv2d(num_channels, nf(res - 1), kernel_size=1, gain=gain,
                                            use_wscale=use_wscale))
            if conditional:
                r = 2 ** res
                self.embeddings.append(nn.Embedding(n_classes, (num_channels // 2) * r * r))
        if self.conditional:
            self.embeddings.append(nn.Embedding(n_classes, (num_channels // 2) * 4 * 4))
            self.embeddings = nn.ModuleList(self.embeddings)
        self.blocks = nn.ModuleList(blocks)
        self.final_block = DiscriminatorTop(self.mbstd_group_size, self.mbstd_num_features, in_channels=nf(2),
                                            intermediate_channels=nf(2), gain=gain, use_wscale=use_wscale,
                                            activation_layer=act)
        from_rgb.append(EqualizedConv2d(num_channels, nf(2), kernel_size=1, gain=gain, use_wscale=use_wscale))
        self.from_rgb = nn.ModuleList(from_rgb)
        self.temporaryDownsampler = lambda x: _ops.PoolNCHW.apply(x, 2, 0.25)

    def _from_rgb(self, img, j):
        m = self.from_rgb[j]
        if img.shape[1] > 4:
            return m.forward_nhwc(to_nhwc(img))                 # conditional: image + label planes (> 4 channels)
        b_mul = m.b_mul if m.bias is not None else 1.0
        return _ops.RgbIn.apply(img, m.weight, m.bias, m.w_mul, b_mul, False, _ops.precision())

    def _embed(self, images_in, labels_in, idx):
        e = self.embeddings[idx](labels_in)
        e = e.view(images_in.shape[0], -1, images_in.shape[2], images_in.shape[3])
        return torch.cat([images_in, e], dim=1)

    def forward(self, images_in, depth, alpha=1., labels_in=None):
        """images [B, C, R, R] float32 NCHW at the resolution of `depth` -> scores [B, 1]."""
        assert depth < self.depth, "Requested output depth cannot be produced"
        _ops._require_cuda(images_in)
        if self.conditional:
            assert labels_in is not None, "Conditional Discriminator requires labels"
        if self.structure == 'fixed':
            if self.conditional:
                images_in = self._embed(images_in, labels_in, 0)
            x = self._from_rgb(images_in, 0)
            for block in self.blocks:
                x = block.forward_nhwc(x)
        elif self.structure == 'linear':
            if depth > 0:
                if self.conditional:
                    images_in = self._embed(images_in, labels_in, self.depth - depth - 1)
                residual = self._from_rgb(self.temporaryDownsampler(images_in), self.depth - depth)
                straight = self.blocks[self.depth - depth - 1].forward_nhwc(
                    self._from_rgb(images_in, self.depth - depth - 1))
                x = _ops.Axpby.apply(straight, residual, *_ops.lerp_coefs(alpha))
                for block in self.blocks[(self.depth - depth):]:
                    x = block.forward_nhwc(x)
            else:
                if self.conditional:
                    images_in = self._embed(images_in, labels_in, -1)
                x = self._from_rgb(images_in, -1)
        else:
            raise KeyError("Unknown structure: ", self.structure)
        return self.final_block.forward_nhwc(x)


class StyleGAN:
    """Wrapper around the Generator and the Discriminator (reference GAN.py:447-826)."""

    def __init__(self, structure, resolution, num_channels, latent_size, g_args, d_args, g_opt_args, d_opt_args,
                 conditional=False, n_classes=0, loss="relativistic-hinge", drift=0.001, d_repeats=1, use_ema=False,
                 ema_decay=0.999, device=torch.device("cpu")):
        assert structure in ['fixed', 'linear']
        if conditional:
            assert n_classes > 0, "Conditional GANs require n_classes > 0"
        self.structure = structure
        self.depth = int(np.log2(resolution)) - 1
        self.latent_size = latent_size
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sgb200: StyleGAN runs on a CUDA device only (got %s); the reference's CPU path has "
                               "no counterpart here" % (self.device,))
        self.d_repeats = d_repeats
        self.use_cuda_graphs = os.environ.get("SG_CUDA_GRAPHS", "1") != "0"
        self.conditional = conditional
        self.n_classes = n_classes
        self.use_ema = use_ema
        self.ema_decay = ema_decay

        self.gen = Generator(num_channels=num_channels, resolution=resolution, structure=self.structure,
                             conditional=self.conditional, n_classes=self.n_classes, **g_args).to(self.device)
        self.dis = Discriminator(num_channels=num_channels, resolution=resolution, structure=self.structure,
                                 conditional=self.conditional, n_classes=self.n_classes, **d_args).to(self.device)

        from . import _ddp
        self.ddp = _ddp.DataParallelContext.from_env(self.device)
        if self.ddp is not None:
            self.ddp.broadcast_module(self.gen)
            self.ddp.broadcast_module(self.dis)

        self.__setup_gen_optim(**g_opt_args)
        self.__setup_dis_optim(**d_opt_args)
        self.drift = drift
        self.loss = self.__setup_loss(loss)
        from . import _graphs
        self._graphs = _graphs.StepGraphs(self)

        if self.use_ema:
            self.gen_shadow = copy.deepcopy(self.gen)
            self.ema_updater = update_average
            self.ema_updater(self.gen_shadow, self.gen, beta=0)
        if self.ddp is not None:
            # after the shadow copy: gen_shadow is evaluated by rank 0 alone (sample grids) and must not join collectives
            self.gen._avg_sync = self.ddp.average_

    def __setup_gen_optim(self, learning_rate, beta_1, beta_2, eps):
        self.gen_optim = FusedAdam(self.gen.parameters(), lr=learning_rate, betas=(float(beta_1), float(beta_2)),
                                   eps=eps)

    def __setup_dis_optim(self, learning_rate, beta_1, beta_2, eps):
        self.dis_optim = FusedAdam(self.dis.parameters(), lr=learning_rate, betas=(float(beta_1), float(beta_2)),
                                   eps=eps)

    def __setup_loss(self, loss):
        if isinstance(loss, str):
            loss = loss.lower()
            if not self.conditional:
                assert loss in ["logistic", "hinge", "standard-gan", "relativistic-hinge"], "Unknown loss function"
                loss_func = {"logistic": Losses.LogisticGAN, "hinge": Losses.HingeGAN,
                             "standard-gan": Losses.StandardGAN,
                             "relativistic-hinge": Losses.RelativisticAverageHingeGAN}[loss](self.dis)
            else:
                assert loss in ["conditional-loss"]
                loss_func = Losses.ConditionalGANLoss(self.dis)
            return loss_func
        return loss

    def __progressive_down_sampling(self, real_batch, depth, alpha):
        """Real images at the resolution of `depth`, blended with the coarser level (reference GAN.py:557-589)."""
        if self.structure == 'fixed':
            return real_batch
        real_batch = real_batch.float()
        f = int(np.power(2, self.depth - depth - 1))
        pf = max(int(np.power(2, self.depth - depth)), 0)
        ds = _ops.PoolNCHW.apply(real_batch, f, 1.0 / (f * f)) if f > 1 else real_batch
        if depth > 0:
            prior = _ops.UpsampleNCHW.apply(_ops.PoolNCHW.apply(real_batch, pf, 1.0 / (pf * pf)), 2, 1.0)
        else:
            prior = ds
        return _ops.Axpby.apply(ds.contiguous(), prior.contiguous(), *_ops.lerp_coefs(alpha))

    def optimize_discriminator(self, noise, real_batch, depth, alpha, labels=None):
        """One discriminator update (reference GAN.py:591-622).  Returns the loss value (host float).
        After two eager calls per (depth, batch shape) the update is replayed from a captured CUDA graph
        (models/_graphs.py; `self.use_cuda_graphs = False` or SG_CUDA_GRAPHS=0 keeps it eager)."""
        loss = self._graphs.run("D", self._d_step, noise, real_batch, depth, alpha, labels)
        return loss.item() / self.d_repeats

    def _d_step(self, noise, real_batch, depth, alpha, labels=None):
        """Device-side part of optimize_discriminator: -> sum of the d_repeats loss values (0-d device tensor)."""
        with torch.no_grad():
            real_samples = self.__progressive_down_sampling(real_batch, depth, alpha)
        loss_sum = None
        for _ in range(self.d_repeats):
            with torch.no_grad():                              # the reference builds, then detaches, this graph
                fake_samples = self.gen(noise, depth, alpha, labels).detach()
            if not self.conditional:
                loss = self.loss.dis_loss(real_samples, fake_samples, depth, alpha)
            else:
                loss = self.loss.dis_loss(real_samples, fake_samples, labels, depth, alpha)
            self.dis_optim.zero_grad()
            loss.backward()
            gscale = self.ddp.allreduce_grads(self.dis_optim) if self.ddp is not None else 1.0
            self.dis_optim.step(grad_scale=gscale)
            loss_sum = loss.detach() if loss_sum is None else loss_sum + loss.detach()
        return loss_sum

    def optimize_generator(self, noise, real_batch, depth, alpha, labels=None):
        """One generator update incl. gradient clipping and EMA (reference GAN.py:624-659)."""
        return self._graphs.run("G", self._g_step, noise, real_batch, depth, alpha, labels).item()

    def _g_step(self, noise, real_batch, depth, alpha, labels=None):
        """Device-side part of optimize_generator: -> loss (0-d device tensor)."""
        real_samples = None
        # the built-in losses below ignore the reals in gen_loss (reference Losses.py:84,128,151,227), so the pyramid
        # the reference always computes (GAN.py:636) is skipped for them only; any other loss object receives it
        ignores_reals = type(self.loss) in (Losses.LogisticGAN, Losses.HingeGAN, Losses.StandardGAN,
                                            Losses.ConditionalGANLoss)
        if not ignores_reals:
            with torch.no_grad():
                real_samples = self.__progressive_down_sampling(real_batch, depth, alpha)
        fake_samples = self.gen(noise, depth, alpha, labels)
        # the discriminator is a fixed function in this step: with requires_grad off its layers compute input
        # gradients only (the reference computes, and then discards, all discriminator weight gradients here)
        d_params = [p for p in self.dis.parameters() if p.requires_grad]
        for p in d_params:
            p.requires_grad_(False)
        try:
            if not self.conditional:
                loss = self.loss.gen_loss(real_samples, fake_samples, depth, alpha)
            else:
                loss = self.loss.gen_loss(real_samples, fake_samples, labels, depth, alpha)
        finally:
            for p in d_params:
                p.requires_grad_(True)
        self.gen_optim.zero_grad()
        # only generator parameters need gradients here (the reference also accumulates unused D gradients)
        g_params = [p for p in self.gen.parameters() if p.requires_grad]
        loss.backward(inputs=g_params)
        gscale = self.ddp.allreduce_grads(self.gen_optim) if self.ddp is not None else 1.0
        # clip_grad_norm_(max_norm=10) + Adam + EMA in one multi-tensor pass
        shadow = dict(self.gen_shadow.named_parameters()) if self.use_ema else None
        ema_map = None
        if shadow is not None:
            ema_map = {p: shadow[n] for n, p in self.gen.named_parameters()}
        self.gen_optim.step(max_norm=10., ema_map=ema_map, ema_decay=self.ema_decay, grad_scale=gscale)
        return loss.detach()

    @staticmethod
    def create_grid(samples, scale_factor, img_file):
        """Save a grid of samples (reference GAN.py:661-680)."""
        from torch.nn.functional import interpolate
        from torchvision.utils import save_image
        samples = samples.detach().float().cpu()
        if scale_factor > 1:
            samples = interpolate(samples, scale_factor=scale_factor)
        save_image(samples, img_file, nrow=int(np.sqrt(len(samples))), normalize=True, scale_each=True,
                   pad_value=128, padding=1)

    def train(self, dataset, num_workers, epochs, batch_sizes, fade_in_percentage, logger, output, num_samples=36,
              start_depth=0, feedback_factor=100, checkpoint_factor=1, resume=None):
        """Progressive training loop (reference GAN.py:682-826): same schedule, logging lines, sample grids and
        checkpoint file names.  Additions that leave the reference behaviour untouched when unused: under data
        parallelism every rank iterates its own shard (`_ddp.sharded_loader`) and only rank 0 writes files; each
        checkpoint gets an extra `GAN_STATE_<depth>_<epoch>.pth` (`_train_state`), and `resume=` (that file or its
        dict) continues at the recorded epoch with the recorded fade-in ticker, step count and host RNG."""
        from data import get_data_loader
        from . import _ddp, _train_state
        assert self.depth <= len(epochs), "epochs not compatible with depth"
        assert self.depth <= len(batch_sizes), "batch_sizes not compatible with depth"
        assert self.depth <= len(fade_in_percentage), "fade_in_percentage not compatible with depth"
        self.gen.train()
        self.dis.train()
        if self.use_ema:
            self.gen_shadow.train()
        is_main = self.ddp is None or self.ddp.rank == 0
        global_time = time.time()
        fixed_input = torch.randn(num_samples, self.latent_size).to(self.device)
        fixed_labels = None
        if self.conditional:
            fixed_labels = torch.linspace(0, self.n_classes - 1, num_samples).to(torch.int64).to(self.device)
        logger.info("Starting the training process ... \n")
        if self.structure == 'fixed':
            start_depth = self.depth - 1
        step = 1
        first_epoch, first_ticker = 1, 1
        if self.ddp is not None and resume is None:
            # torch's default CPU seed is the same constant in every process: without this all ranks would draw the
            # same latents, style-mixing latents and NoiseLayer noise, i.e. `world` copies of one fake batch.
            # fixed_input was drawn above, before the streams separate, so the sample grids do not depend on the rank.
            from . import CustomLayers
            base = torch.initial_seed() % (2 ** 31)
            torch.manual_seed(base + self.ddp.rank)
            random.seed(base + self.ddp.rank)
            np.random.seed((base + self.ddp.rank) % (2 ** 32))
            CustomLayers.manual_seed(base + self.ddp.rank)
        if resume is not None:
            state = _train_state.load(resume)
            start_depth, first_epoch, first_ticker, step = _train_state.resume_point(state, epochs)
            _train_state.restore_rng(state, 0 if self.ddp is None else self.ddp.rank)
            logger.info("Resuming at depth %d, epoch %d, step %d", start_depth + 1, first_epoch, step)
        for current_depth in range(start_depth, self.depth):
            current_res = np.power(2, current_depth + 2)
            logger.info("Currently working on depth: %d", current_depth + 1)
            logger.info("Current resolution: %d x %d" % (current_res, current_res))
            resumed_depth = resume is not None and current_depth == start_depth
            ticker = first_ticker if resumed_depth else 1
            if self.ddp is None:
                data = get_data_loader(dataset, batch_sizes[current_depth], num_workers)
            else:
                data = _ddp.sharded_loader(dataset, batch_sizes[current_depth], num_workers, self.ddp.rank,
                                           self.ddp.world)
            for epoch in range(first_epoch if resumed_depth else 1, epochs[current_depth] + 1):
                start = timeit.default_timer()
                logger.info("Epoch: [%d]" % epoch)
                if self.ddp is not None:
                    data.sampler.set_epoch(current_depth * 1000 + epoch)
                total_batches = len(data)
                fade_point = int((fade_in_percentage[current_depth] / 100) * epochs[current_depth] * total_batches)
                for i, batch in enumerate(data, 1):
                    alpha = ticker / fade_point if ticker <= fade_point else 1
                    if self.conditional:
                        images, labels = batch
                        labels = labels.to(self.device)
                    else:
                        images, labels = batch, None
                    images = images.to(self.device)
                    gan_input = torch.randn(images.shape[0], self.latent_size).to(self.device)
                    dis_loss = self.optimize_discriminator(gan_input, images, current_depth, alpha, labels)
                    gen_loss = self.optimize_generator(gan_input, images, current_depth, alpha, labels)
                    if i % int(total_batches / feedback_factor + 1) == 0 or i == 1:
                        elapsed = str(datetime.timedelta(seconds=time.time() - global_time)).split('.')[0]
                        logger.info("Elapsed: [%s] Step: %d  Batch: %d  D_Loss: %f  G_Loss: %f"
                                    % (elapsed, step, i, dis_loss, gen_loss))
                        if is_main:
                            os.makedirs(os.path.join(output, 'samples'), exist_ok=True)
                            gen_img_file = os.path.join(output, 'samples', "gen_" + str(current_depth) + "_"
                                                        + str(epoch) + "_" + str(i) + ".png")
                            with torch.no_grad():
                                net = self.gen_shadow if self.use_ema else self.gen
                                self.create_grid(
                                    samples=net(fixed_input, current_depth, alpha, labels_in=fixed_labels).detach(),
                                    scale_factor=int(np.power(2, self.depth - current_depth - 1))
                                    if self.structure == 'linear' else 1,
                                    img_file=gen_img_file)
                    ticker += 1
                    step += 1
                elapsed = str(datetime.timedelta(seconds=timeit.default_timer() - start)).split('.')[0]
                logger.info("Time taken for epoch: %s\n" % elapsed)
                if is_main and (epoch % checkpoint_factor == 0 or epoch == 1 or epoch == epochs[current_depth]):
                    save_dir = os.path.join(output, 'models')
                    os.makedirs(save_dir, exist_ok=True)
                    tag = str(current_depth) + "_" + str(epoch) + ".pth"
                    gen_save_file = os.path.join(save_dir, "GAN_GEN_" + tag)
                    torch.save(self.gen.state_dict(), gen_save_file)
                    logger.info("Saving the model to: %s\n" % gen_save_file)
                    torch.save(self.dis.state_dict(), os.path.join(save_dir, "GAN_DIS_" + tag))
                    torch.save(self.gen_optim.state_dict(), os.path.join(save_dir, "GAN_GEN_OPTIM_" + tag))
                    torch.save(self.dis_optim.state_dict(), os.path.join(save_dir, "GAN_DIS_OPTIM_" + tag))
                    if self.use_ema:
                        gen_shadow_save_file = os.path.join(save_dir, "GAN_GEN_SHADOW_" + tag)
                        torch.save(self.gen_shadow.state_dict(), gen_shadow_save_file)
                        logger.info("Saving the model to: %s\n" % gen_shadow_save_file)
                    torch.save(_train_state.capture(current_depth, epoch, ticker, step),
                               os.path.join(save_dir, "GAN_STATE_" + tag))
        logger.info('Training completed.\n')


if __name__ == '__main__':
    print('Done.')
