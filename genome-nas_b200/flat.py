"""Flat fp32 storage for parameters, gradients and BatchNorm running statistics.

The reference keeps one nn.Parameter per weight tensor and its state_dict keys
are a contract (CWP-DARTS / OSP-NAS copy weights by key).  The kernels want raw
device pointers, so every Parameter stays an nn.Parameter but is made a *view*
into one contiguous buffer per module tree; the C descriptors then carry only
offsets relative to the first tensor.  Gradients mirror the layout (p.grad are
views of one buffer), which is also what the NCCL allreduce and the fused
optimiser kernels consume.  Every tensor starts at a multiple of 4 floats so the
kernels may use 128-bit loads.  A region adopts the layout of an enclosing
region (a cell inside the supernet) instead of re-allocating.
"""
from typing import Dict, List, Tuple

import torch

ALIGN = 4
_BN_SUFFIX = ("running_mean", "running_var")


def _aligned(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


def layout_of(named) -> Tuple[Dict[str, int], int]:
    off, out = 0, {}
    for name, t in named:
        out[name] = off
        off += _aligned(t.numel())
    return out, off


def _is_flat(tensors, offs):
    base, dev = tensors[0].data_ptr(), tensors[0].device
    for t, o in zip(tensors, offs):
        if t.dtype != torch.float32 or t.device != dev or not t.is_contiguous() or t.data_ptr() != base + 4 * o:
            return False
    return True


def _alias(t, total):
    """1-D view of `total` floats starting at t's first element, if t's storage is large enough."""
    if t.untyped_storage().nbytes() // 4 - t.storage_offset() < total:
        return None
    return torch.as_strided(t.detach(), (total,), (1,), t.storage_offset())


class FlatRegion:
    """Contiguous storage for one module tree (a MixedOp, a cell, an RHN layer or the whole supernet)."""

    def __init__(self, module, exclude=()):
        self.module = module
        self.exclude = tuple(exclude)   # parameter names owned by an enclosing module (e.g. the shared RNN alphas)
        self.flat = None      # parameters
        self.gflat = None     # gradients
        self.bn = None        # running_mean | running_var of every BatchNorm, in registration order
        self.offsets: Dict[str, int] = {}
        self.bn_offsets: Dict[str, int] = {}
        self.total = 0
        self._names: List[str] = []
        self._params: List[torch.nn.Parameter] = []
        self._offs: List[int] = []
        self._probe = None
        self._gviews = None

    def _bufs(self):
        return [(n, b) for n, b in self.module.named_buffers() if n.endswith(_BN_SUFFIX)]

    def _make_probe(self, bufs):
        p = self._params
        probe = (p[0].device, len(p)) + tuple(q.data_ptr() for q in p[::61]) + (p[-1].data_ptr(),) if p else ()
        if bufs:
            probe += (bufs[0][1].data_ptr(), bufs[-1][1].data_ptr(), bufs[0][1].device)
        return probe

    def ensure(self, device=None):
        named = [(n, p) for n, p in self.module.named_parameters() if n not in self.exclude]
        self._names = [n for n, _ in named]
        self._params = [p for _, p in named]
        bufs = self._bufs()
        if not named and not bufs:      # e.g. a MixedOp reduced to {none, skip_connect}: nothing to lay out
            if self.flat is None or (device is not None and self.flat.device != torch.device(device)):
                self.flat = torch.zeros(4, dtype=torch.float32, device=device)
                self.bn = torch.zeros(4, dtype=torch.float32, device=device)
                self.gflat = None
            return self
        if self._probe is not None and self._make_probe(bufs) == self._probe:
            return self
        for p in self._params:
            if p.dtype != torch.float32:
                raise RuntimeError("genome_nas_b200 kernels are fp32: call .float() on the module")
        self.offsets, self.total = layout_of(named)
        self._offs = [self.offsets[n] for n in self._names]
        dev = self._params[0].device if self._params else bufs[0][1].device
        datas = [p.data for p in self._params]
        flat = _alias(datas[0], self.total) if (datas and _is_flat(datas, self._offs)) else None
        if flat is None:
            flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for p, o in zip(self._params, self._offs):
                    flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                    p.data = flat[o:o + p.numel()].view(p.shape)
        self.flat = flat
        self._gviews = None
        if self.gflat is not None and (self.gflat.numel() != self.total or self.gflat.device != dev):
            self.gflat = None
        if bufs:
            self.bn_offsets, btotal = layout_of(bufs)
            btens = [b for _, b in bufs]
            boffs = [self.bn_offsets[n] for n, _ in bufs]
            bn = _alias(btens[0], btotal) if _is_flat(btens, boffs) else None
            if bn is None:
                bn = torch.zeros(btotal, dtype=torch.float32, device=dev)
                with torch.no_grad():
                    for (n, b), o in zip(bufs, boffs):
                        bn[o:o + b.numel()].copy_(b.reshape(-1))
                        b.set_(bn[o:o + b.numel()].view(b.shape))
            self.bn = bn
        self._probe = self._make_probe(bufs)
        return self

    # -- gradients -------------------------------------------------------
    def ensure_grads(self):
        """Make every p.grad a view of one flat buffer (zero-filled where it was unset); return the buffer."""
        self.ensure(self.flat.device if self.flat is not None else None)
        if not self._params:
            if self.gflat is None:
                self.gflat = torch.zeros(4, dtype=torch.float32, device=self.flat.device)
            return self.gflat
        grads = [p.grad for p in self._params]
        n_unset = sum(g is None for g in grads)
        if n_unset == 0 and _is_flat(grads, self._offs):
            if self.gflat is None or self.gflat.data_ptr() != grads[0].data_ptr():
                g = _alias(grads[0], self.total)
                if g is not None:
                    self.gflat, self._gviews = g, None
                    return self.gflat
            else:
                return self.gflat
        if self.gflat is None:
            self.gflat = torch.zeros(self.total, dtype=torch.float32, device=self.flat.device)
            self._gviews = None
        if self._gviews is None:
            self._gviews = [self.gflat[o:o + p.numel()].view(p.shape) for p, o in zip(self._params, self._offs)]
        if n_unset == len(grads):
            self.gflat.zero_()
            for p, v in zip(self._params, self._gviews):
                p.grad = v
            return self.gflat
        gbase = self.gflat.data_ptr()
        with torch.no_grad():
            for p, v, o, g in zip(self._params, self._gviews, self._offs, grads):
                if g is None:
                    v.zero_()
                    p.grad = v
                elif g.data_ptr() != gbase + 4 * o:
                    v.copy_(g)
                    p.grad = v
        return self.gflat

    def zero_grads(self):
        """Fast replacement for optimizer.zero_grad(): one memset, the p.grad views stay attached."""
        self.ensure_grads()
        self.gflat.zero_()


def region_of(module, exclude=(), device=None) -> FlatRegion:
    r = module.__dict__.get("_gnas_region")
    if r is None:
        r = FlatRegion(module, exclude)
        module.__dict__["_gnas_region"] = r
    return r.ensure(device)
