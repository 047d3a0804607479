"""Resumable position of `StyleGAN.train`: the reference checkpoints the five state dicts per epoch
(`models/GAN.py:804-823`) but not where in the progressive schedule it was, so a restart begins the depth again with
alpha = 0.  `GAN_STATE_<depth>_<epoch>.pth` is an EXTRA file beside the reference's five (their names and contents are
unchanged): schedule position, fade-in ticker, global step and the host RNG states the step consumes
(`torch.randn` latents on the CPU generator, `random` for style mixing -- GAN.py:282-286,765)."""
import random

import numpy as np
import torch

VERSION = 1


def capture(depth, epoch, ticker, step):
    """State after finishing `epoch` (1-based) of `depth`: training resumes at epoch + 1 with the same ticker."""
    from .CustomLayers import _Noise
    return {"version": VERSION, "depth": int(depth), "epoch": int(epoch), "ticker": int(ticker), "step": int(step),
            "torch_rng": torch.get_rng_state(), "python_rng": random.getstate(), "numpy_rng": np.random.get_state(),
            "noise_rng": (_Noise.seed, _Noise.offset)}     # the engine's Philox stream of NoiseLayer


def restore_rng(state, rank=0):
    """Rank 0 (and single-process runs) continue the recorded random streams exactly.  The file is written by rank 0
    only, so the other ranks of a data-parallel job take fresh streams derived from (step, rank): they must not
    repeat rank 0's latents."""
    from .CustomLayers import _Noise
    if rank == 0:
        torch.set_rng_state(state["torch_rng"])
        random.setstate(state["python_rng"])
        np.random.set_state(state["numpy_rng"])
        if state.get("noise_rng") is not None:
            _Noise.seed, _Noise.offset = state["noise_rng"]
    else:
        seed = (int(state["step"]) * 1000003 + int(rank) * 7919 + 12345) % (2 ** 31 - 1)
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        _Noise.seed, _Noise.offset = seed, 0


def resume_point(state, epochs):
    """-> (start_depth, first_epoch, ticker, step) for a schedule with `epochs[d]` epochs at depth d."""
    if state.get("version") != VERSION:
        raise ValueError("unknown train-state version %r" % (state.get("version"),))
    depth, epoch = state["depth"], state["epoch"]
    if epoch >= epochs[depth]:                      # the depth was finished: start the next one from alpha = 0
        return depth + 1, 1, 1, state["step"]
    return depth, epoch + 1, state["ticker"], state["step"]


def load(path_or_state):
    if isinstance(path_or_state, dict):
        return path_or_state
    return torch.load(path_or_state, map_location="cpu", weights_only=False)
