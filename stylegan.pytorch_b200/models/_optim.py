"""Fused multi-tensor Adam (+ global-norm clip, + 1/world gradient scale, + generator EMA) on the sgb200 kernels.

``FusedAdam`` subclasses ``torch.optim.Adam`` so that ``state_dict()`` / ``load_state_dict()`` keep the exact
layout of the reference's GAN_*_OPTIM_*.pth files (reference GAN.py:529-533, 816-817): per-parameter ``step``,
``exp_avg``, ``exp_avg_sq`` in ``module.parameters()`` order.  ``step()`` is one kernel launch over all tensors
(plus a one-block step-counter bump) instead of hundreds of per-tensor ops."""
import torch

from . import _ops
from ._lib import lib, ptr, stream, count_launch


def _table(entries, device):
    """entries: list of (p, g_ptr, m, v, ema_or_None, step, n) -> (int64 [T,8] device tensor, total_chunks)."""
    chunk = lib.sg_chunk_elems()
    rows, c0 = [], 0
    for p_ptr, g_ptr, m_ptr, v_ptr, e_ptr, s_ptr, n in entries:
        rows.append([p_ptr, g_ptr, m_ptr, v_ptr, e_ptr, s_ptr, n, c0])
        c0 += (n + chunk - 1) // chunk
    t = torch.tensor(rows, dtype=torch.int64).reshape(-1, 8)
    return _ops.upload_table(t, device), c0


_EMA_TOKEN = [0]


def ema_update(pairs, beta):
    """tgt = beta*tgt + (1-beta)*src for every (tgt, src) parameter pair, one launch."""
    pairs = [(t, s) for t, s in pairs if t.numel() > 0]
    for t, _ in pairs:                       # the averaged copies have their own update counter (derived-weight cache)
        if getattr(t, "_sg_token", None) is not _EMA_TOKEN:
            t._sg_token = _EMA_TOKEN
    if not pairs:
        return
    dev = pairs[0][0].device
    _ops._require_cuda(pairs[0][0])
    entries = [(s.data_ptr(), 0, 0, 0, t.data_ptr(), 0, t.numel()) for t, s in pairs]
    tab, chunks = _table(entries, dev)
    lib.sg_ema_multi(ptr(tab), len(entries), chunks, float(beta), stream())
    count_launch()
    _EMA_TOKEN[0] += 1                       # variants of the averaged weights are rebuilt lazily, on their next use


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=(float(betas[0]), float(betas[1])), eps=eps)
        self._token = [0]              # update counter shared by this optimizer's parameters (_ops.cached_weight)
        for g in self.param_groups:
            g["capturable"] = True     # keeps state['step'] on the parameter's device through load_state_dict
            for p in g["params"]:
                p._sg_token = self._token

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif st["step"].device != p.device or st["step"].dtype != torch.float32:
            st["step"] = st["step"].to(device=p.device, dtype=torch.float32)
        return st

    @torch.no_grad()
    def step(self, closure=None, max_norm=None, ema_map=None, ema_decay=0.999, grad_scale=1.0):
        """max_norm: clip_grad_norm_ threshold applied to the (scaled) gradients of THIS optimizer's parameters.
        ema_map: {param: shadow_param}; updated in the same pass (parameters without a gradient are averaged in a
        second launch, as models/__init__.py:update_average touches every parameter).
        grad_scale: multiplies every gradient (1/world after a SUM all-reduce, models/_ddp.py)."""
        assert closure is None
        for group in self.param_groups:
            lr, (b1, b2), eps = group["lr"], group["betas"], group["eps"]
            entries, rest = [], []
            dev = None
            for p in group["params"]:
                g = p.grad
                if g is None:
                    if ema_map is not None and p in ema_map:
                        rest.append((ema_map[p], p))
                    continue
                _ops._require_cuda(p)
                dev = p.device
                st = self._state_for(p)
                if not g.is_contiguous() or g.dtype != torch.float32:
                    g = g.contiguous().float()
                    p.grad = g
                g_ptr = g.data_ptr()
                e_ptr = ema_map[p].data_ptr() if (ema_map is not None and p in ema_map) else 0
                entries.append((p.data_ptr(), g_ptr, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), e_ptr,
                                st["step"].data_ptr(), p.numel()))
            if entries:
                tab, chunks = _table(entries, dev)
                normsq = None
                if max_norm is not None:
                    nbuf = torch.empty((1 + chunks,), dtype=torch.float64, device=dev)
                    normsq = nbuf[:1]
                    lib.sg_sqnorm_multi(ptr(tab), len(entries), chunks, ptr(normsq), nbuf.data_ptr() + 8, stream())
                    count_launch()
                lib.sg_adam_multi(ptr(tab), len(entries), chunks, float(lr), float(b1), float(b2), float(eps),
                                  float(grad_scale), ptr(normsq), float(max_norm or 0.0), float(ema_decay), stream())
                count_launch(2)
                self._keepalive = (tab, normsq)
            if rest:
                ema_update(rest, ema_decay)
            if entries:
                # the parameters changed under autograd's feet: refresh their derived (scaled / packed / bf16) copies in
                # place, here, so that the next forward -- eager or replayed from a CUDA graph -- finds them current
                updated = [p for p in group["params"] if p.grad is not None]
                self._token[0] += 1
                if ema_map is not None:
                    _EMA_TOKEN[0] += 1
                    for p in updated:
                        if p in ema_map and getattr(ema_map[p], "_sg_token", None) is not _EMA_TOKEN:
                            ema_map[p]._sg_token = _EMA_TOKEN
                _ops.refresh_derived(updated)
        return None
