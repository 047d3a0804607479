"""Drop-in replacement for the reference's ``models`` package (huangzh13/StyleGAN.pytorch, models/__init__.py).

Same import surface -- ``models.GAN.{StyleGAN, Generator, Discriminator}``, ``models.Blocks``,
``models.CustomLayers``, ``models.Losses``, ``models.update_average`` -- but the per-step hot path executes in
libsgb200.so (hand-written sm_100a CUDA, see ../csrc).  Engine controls: ``models.set_precision('fp32'|'bf16')``."""


def update_average(model_tgt, model_src, beta):
    """EMA of generator parameters into the shadow copy (reference models/__init__.py:13-40): parameters only,
    buffers untouched, matched by name.  Runs as one multi-tensor kernel launch."""
    import torch
    from ._optim import ema_update
    src = dict(model_src.named_parameters())
    pairs = []
    for name, p_tgt in model_tgt.named_parameters():
        p_src = src[name]
        assert p_src is not p_tgt
        pairs.append((p_tgt, p_src))
    with torch.no_grad():
        ema_update(pairs, beta)


def set_precision(p):
    from ._ops import set_precision as _sp
    _sp(p)


def precision():
    from ._ops import precision as _p
    return _p()


def set_shared_r1(flag):
    """Evaluate D(real) once per discriminator step for both the logistic term and the R1 penalty (see
    Losses.LogisticGAN.shared_real_pass); default from the environment variable SG_SHARED_R1."""
    from .Losses import LogisticGAN
    LogisticGAN.shared_real_pass = bool(flag)
