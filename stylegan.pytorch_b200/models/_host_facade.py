"""Transparent device placement for the inference scripts.

`generate_samples.py`, `generate_mixing_figure.py` and `generate_truncation_figure.py` build a `Generator` on the CPU,
load a checkpoint into it and call `gen(...)`, `gen.g_mapping(...)`, `gen.g_synthesis(...)` with CPU tensors, then use
`.numpy()` on the results (`generate_samples.py:95-110`, `generate_mixing_figure.py:12-48`).  There is no CPU
implementation of the kernels, so such a call is executed on the GPU through a cached device replica of the module and
its results are returned as CPU tensors; buffers the call updates (`truncation.avg_latent`) are copied back.  Nothing
here runs when the caller already works with CUDA tensors, and without a CUDA device the call still fails loudly."""
import copy

import torch

_KEY = "_sg_device_replica"


def wants_facade(module, *tensors):
    """True when a CPU-resident module is called with CPU tensors and a GPU exists to run it on."""
    seen = False
    for t in tensors:                                  # the training step leaves here on its first (CUDA) tensor
        if isinstance(t, torch.Tensor):
            if t.device.type != "cpu":
                return False
            seen = True
    if not seen or not torch.cuda.is_available():
        return False
    p = next(module.parameters(), None)
    return p is not None and p.device.type == "cpu"


def _version_key(module):
    return (module.training,) + tuple(t._version for t in module.parameters()) + \
        tuple(t._version for t in module.buffers()) + tuple(id(t) for t in module.parameters())


def _replica(module):
    cached = module.__dict__.get(_KEY)
    key = _version_key(module)
    if cached is not None and cached[0] == key:
        return cached[1]
    for m in module.modules():                        # never deep-copy an old replica along with the module
        m.__dict__.pop(_KEY, None)
    rep = copy.deepcopy(module).to(torch.device("cuda", torch.cuda.current_device()))
    rep.train(module.training)
    module.__dict__[_KEY] = (key, rep)
    return rep


def _to_device(x, dev):
    return x.to(dev) if isinstance(x, torch.Tensor) else x


def _to_host(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu() if not x.requires_grad else x.cpu()
    if isinstance(x, (tuple, list)):
        return type(x)(_to_host(v) for v in x)
    return x


def call(module, *args, **kwargs):
    rep = _replica(module)
    dev = next(rep.parameters()).device
    out = rep(*[_to_device(a, dev) for a in args], **{k: _to_device(v, dev) for k, v in kwargs.items()})
    # buffers the forward updated on the device (the W average of training-mode calls) go back to the host module
    changed = False
    host_buffers = dict(module.named_buffers())
    with torch.no_grad():
        for name, buf in rep.named_buffers():
            hb = host_buffers.get(name)
            if hb is not None and not torch.equal(hb, buf.detach().cpu()):
                hb.copy_(buf.detach().cpu())
                changed = True
    if changed:                                        # keep the replica valid for the next call
        module.__dict__[_KEY] = (_version_key(module), rep)
    return _to_host(out)
