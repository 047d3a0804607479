"""CUDA-graph replay of the two optimizer steps (StyleGAN.optimize_discriminator / optimize_generator).

A G+D step at 256x256 is ~1 100 kernel launches and ~600 autograd Function calls: the Python + launch cost of driving
them (~32 ms, measured with a batch small enough to leave the GPU idle) is of the same size as the GPU work itself
(~37 ms at batch 32).  After two eager calls per (step kind, depth, shapes) the device-side part of the step --
forward, the R1 double backward, backward, gradient all-reduce, clip + Adam + EMA -- is captured once with
torch.cuda.graph and replayed; the reference's per-step host semantics are kept by feeding everything that changes from
step to step through device memory instead of kernel arguments:

  * latents and images: copied into the graph's static input buffers;
  * the fade-in alpha (reference GAN.py:202, 427, 588): a device (alpha, 1 - alpha) pair read by sg_axpby_dev;
  * style mixing (GAN.py:282-290): the second latents and the cut-off are drawn on the host by the same calls in the
    same order as the eager path (torch.randn on the CPU generator, random.random(), random.randint()) and copied in;
  * NoiseLayer noise (CustomLayers.py:193): the Philox counter base of the step (sg_randn_dev);
  * loss value: a device scalar read after the replay.

TMA descriptors are encoded on the host at capture time; they stay valid because all activations of a captured step
live in the graph's private memory pool, which replays the same addresses."""
import random
import os
import warnings

import torch

from . import _ops
from ._lib import count_launch, launches

WARMUP_EAGER_CALLS = 2


class _Capture:
    """Per-graph dynamic inputs + the state the ops consult while a capture is in progress (_ops.capture_ctx())."""

    def __init__(self, device, z_shape, mixing_draws):
        self.keep = []
        self.device = device
        self.ab = torch.zeros(2, dtype=torch.float32, device=device)            # (alpha, 1 - alpha)
        self.ab_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.cut = torch.zeros(max(1, mixing_draws), dtype=torch.int32, device=device)
        self.cut_host = torch.zeros(max(1, mixing_draws), dtype=torch.int32).pin_memory()
        self.rng_base = torch.zeros(1, dtype=torch.int64, device=device)
        self.rng_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.lat2 = [torch.zeros(z_shape, dtype=torch.float32, device=device) for _ in range(mixing_draws)]
        self.lat2_host = [torch.zeros(z_shape, dtype=torch.float32).pin_memory() for _ in range(mixing_draws)]
        self.rng_offset0 = 0
        self._mix_i = 0
        # pinned staging for the small pointer tables the step uploads (multi-tensor Adam, grouped linears): allocated
        # BEFORE the capture (cudaHostAlloc is not a capturable call) and alive as long as the graph, because the captured
        # copy nodes re-read these host addresses at every replay
        self._arena = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
        self._arena_used = 0

    def pinned_copy(self, t):
        """t: CPU tensor -> a pinned tensor with the same contents carved from the arena"""
        nbytes = (t.numel() * t.element_size() + 63) // 64 * 64
        assert self._arena_used + nbytes <= self._arena.numel(), "sgb200: pinned staging arena exhausted"
        view = self._arena[self._arena_used:self._arena_used + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
        self._arena_used += nbytes
        view.copy_(t)
        return view

    def next_mixing(self, shape):
        """called by Generator.forward during capture: the i-th mixing draw of the step as device-resident values"""
        i = self._mix_i
        self._mix_i += 1
        assert i < len(self.lat2) and tuple(shape) == tuple(self.lat2[i].shape)
        return self.lat2[i], _ops.DynInt(self.cut[i:i + 1], 0)


class _Entry:
    def __init__(self):
        self.eager_calls = 0
        self.graph = None
        self.failed = False


class StepGraphs:
    def __init__(self, sg):
        self.sg = sg
        self.entries = {}
        self.pool = None

    # -- what a Generator.forward in training mode draws from the host RNG ----------------------------------------
    def _mixing_draws(self, kind):
        g = self.sg.gen
        if not g.training or g.style_mixing_prob is None or g.style_mixing_prob <= 0:
            return 0
        return self.sg.d_repeats if kind == "D" else 1

    def _eligible(self, noise, real, labels):
        sg = self.sg
        return (sg.use_cuda_graphs and labels is None and not sg.conditional and noise.is_cuda and real.is_cuda
                and torch.is_grad_enabled() and sg.gen.training and sg.dis.training)

    def run(self, kind, body, noise, real, depth, alpha, labels):
        if not self._eligible(noise, real, labels):
            return body(noise, real, depth, alpha, labels)
        key = (kind, int(depth), tuple(noise.shape), tuple(real.shape), real.dtype, str(_ops.precision()),
               type(self.sg.loss).__name__, getattr(type(self.sg.loss), "shared_real_pass", None))
        e = self.entries.get(key)
        if e is None:
            e = self.entries[key] = _Entry()
        if e.failed:
            return body(noise, real, depth, alpha, labels)
        if e.graph is None:
            if e.eager_calls < WARMUP_EAGER_CALLS:
                e.eager_calls += 1
                return body(noise, real, depth, alpha, labels)
            try:
                self._capture(e, kind, body, noise, real, depth)
            except Exception as ex:                     # stay correct: fall back to the eager path for this key
                e.failed, e.graph = True, None
                _ops._capture[0] = None
                import traceback
                where = " <- ".join("%s:%d %s" % (os.path.basename(f.filename), f.lineno, f.name)
                                    for f in traceback.extract_tb(ex.__traceback__)[-6:])
                warnings.warn("sgb200: CUDA-graph capture of the %s step failed (%s: %s) at %s; running it eagerly"
                              % (kind, type(ex).__name__, str(ex)[:200], where))
                return body(noise, real, depth, alpha, labels)
        return self._replay(e, noise, real, depth, alpha)

    def _capture(self, e, kind, body, noise, real, depth):
        from .CustomLayers import _Noise
        dev = noise.device
        e.z = torch.empty_like(noise)
        e.img = torch.empty_like(real)
        e.cap = cap = _Capture(dev, tuple(noise.shape), self._mixing_draws(kind))
        e.depth = depth
        e.mix_prob = self.sg.gen.style_mixing_prob
        alpha = _ops.DynScalar(cap.ab, 0.5)
        if _Noise.seed is None:
            _Noise.seed = torch.initial_seed()
        cap.rng_offset0 = off0 = _Noise.offset
        l0 = launches()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        _ops._capture[0] = cap
        # host-side state a failed capture must give back (the caller then runs the same step eagerly)
        import random
        host_state = (random.getstate(), torch.random.get_rng_state())
        try:
            # thread_local: CUDA calls of OTHER threads stay legal during the capture -- with torch.distributed the NCCL
            # watchdog thread queries events of earlier collectives, which invalidates a capture in the default global mode
            # (seen with world size 2: cudaErrorStreamCaptureInvalidated)
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                e.loss = body(e.z, e.img, depth, alpha, None)
        except BaseException:
            random.setstate(host_state[0])
            torch.random.set_rng_state(host_state[1])
            _Noise.offset = off0
            count_launch(l0 - launches())
            raise
        finally:
            _ops._capture[0] = None
        if self.pool is None:
            self.pool = g.pool()
        e.launches = launches() - l0
        count_launch(-e.launches)                      # nothing ran during capture
        e.rng_per_step = _Noise.offset - off0
        _Noise.offset = off0
        e.graph = g

    def _replay(self, e, noise, real, depth, alpha):
        from .CustomLayers import _Noise
        from .GAN import draw_style_mixing
        cap = e.cap
        e.z.copy_(noise, non_blocking=True)
        e.img.copy_(real, non_blocking=True)
        a = float(alpha)
        cap.ab_host[0], cap.ab_host[1] = a, 1.0 - a
        cap.ab.copy_(cap.ab_host, non_blocking=True)
        for i in range(len(cap.lat2)):                  # the host draws of the step, in the reference's order
            lat2, cutoff = draw_style_mixing(tuple(e.z.shape), depth, e.mix_prob)
            cap.lat2_host[i].copy_(lat2)
            cap.lat2[i].copy_(cap.lat2_host[i], non_blocking=True)
            cap.cut_host[i] = cutoff
        if len(cap.lat2):
            cap.cut.copy_(cap.cut_host, non_blocking=True)
        cap.rng_host[0] = _Noise.offset
        cap.rng_base.copy_(cap.rng_host, non_blocking=True)
        e.graph.replay()
        _Noise.offset += e.rng_per_step
        count_launch(e.launches)                        # (derived weight copies are refreshed inside the graph, in place)
        return e.loss
