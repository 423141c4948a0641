"""MixedOp and CNN_Cell_search with the reference's module API, evaluated by the
sm_100a kernels through the C ABI (include/gnas.h).

Reference interface mirrored: darts_tools/model_discCNN.py:46-93 (MixedOp) and
:96-201 (CNN_Cell_search) -- constructor arguments, `forward(x, weights)` /
`forward(s0, s1, weights)`, `update_p()`, attributes `m_ops`, `cell_ops`,
`preprocess0/1`, `reduction`, `p`, and every state_dict key.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .flat import region_of
from .operations import (OPS, PRIMITIVES_cnn, FactorizedReduce, Identity, ReLUConvBN)

# parameter / BatchNorm-buffer names of each primitive below "<edge>.m_ops.<primitive>"
_W_NAMES = {
    'sep_conv_15': ('op.1.weight', 'op.2.weight', 'op.5.weight', 'op.6.weight'),
    'sep_conv_9': ('op.1.weight', 'op.2.weight', 'op.5.weight', 'op.6.weight'),
    'dil_conv_15': ('op.1.weight', 'op.2.weight'),
    'dil_conv_9': ('op.1.weight', 'op.2.weight'),
    'conv_15': ('op.1.weight', 'op.4.weight'),
}
_BN_NAMES = {
    'max_pool_5': ('1.running_mean',),
    'avg_pool_5': ('1.running_mean',),
    'sep_conv_15': ('op.3.running_mean', 'op.7.running_mean'),
    'sep_conv_9': ('op.3.running_mean', 'op.7.running_mean'),
    'dil_conv_15': ('op.3.running_mean',),
    'dil_conv_9': ('op.3.running_mean',),
    'conv_15': ('op.2.running_mean', 'op.5.running_mean'),
}


def _fill_edge(desc, region, e, prefix, switch, stride, holder="m_ops", by_index=False):
    """offsets of one edge's parameters / BN buffers; `holder` and `by_index` select the child naming:
    `m_ops.<primitive>` (DARTS family) or `_ops.<index>` (mask supernets)"""
    for k, name in enumerate(PRIMITIVES_cnn):
        desc.sw[e][k] = 1 if switch[k] else 0
        for i in range(4):
            desc.w_off[e][k][i] = -1
        for i in range(2):
            desc.bn_off[e][k][i] = -1
        if not switch[k]:
            continue
        base = f"{prefix}{holder}.{k if by_index else name}."
        if name == 'skip_connect':
            if stride > 1:
                o1, o2 = region.offsets[base + 'conv_1.weight'], region.offsets[base + 'conv_2.weight']
                half = region.module.get_parameter(base + 'conv_1.weight').numel()
                if o2 != o1 + half:
                    raise RuntimeError("FactorizedReduce conv_1/conv_2 must be contiguous in the flat buffer")
                desc.w_off[e][k][0] = o1
                desc.bn_off[e][k][0] = region.bn_offsets[base + 'bn.running_mean']
            continue
        for i, wn in enumerate(_W_NAMES.get(name, ())):
            desc.w_off[e][k][i] = region.offsets[base + wn]
        for i, bn in enumerate(_BN_NAMES.get(name, ())):
            desc.bn_off[e][k][i] = region.bn_offsets[base + bn]


class _FusedCellFunction(torch.autograd.Function):
    """One autograd node per cell (or lone MixedOp).  Weight gradients are written straight into the
    flat gradient buffer the parameters' .grad alias (FlatRegion), mixing-weight and input gradients
    are returned to autograd."""

    @staticmethod
    def forward(ctx, runner, s0, s1, weights, _anchor):
        lib = _lib.lib()
        region = region_of(runner, device=s1.device)
        has_pre = runner._has_pre
        s1 = _lib.require(s1.contiguous(), "s1")
        if has_pre:
            s0 = _lib.require(s0.contiguous(), "s0")
        weights = _lib.require(weights.detach().contiguous(), "weights")
        desc = runner._descriptor(region, s0, s1, weights)
        fwd, bwd = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.gnas_cell_workspace(C.byref(desc), C.byref(fwd), C.byref(bwd)), "cell_workspace")
        dev = s1.device
        ws = torch.empty(fwd.value, dtype=torch.float32, device=dev)
        n_out = 4 if has_pre else 1
        out = torch.empty(desc.B, n_out * desc.C, desc.L_out, dtype=torch.float32, device=dev)
        masks, mask_arr = runner._draw_skip_masks(desc, s1)
        _lib.check(lib.gnas_cell_forward(C.byref(desc), _lib.ptr(region.flat), _lib.ptr(region.bn),
                                         _lib.ptr(s0) if has_pre else None, _lib.ptr(s1), _lib.ptr(weights),
                                         mask_arr, _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr(s1)), "cell_forward")
        _lib.LAUNCHES += 1
        if desc.training:
            runner._bump_batches_tracked()
        ctx.runner, ctx.desc, ctx.masks, ctx.mask_arr = runner, desc, masks, mask_arr
        ctx.bwd_floats = bwd.value
        ctx.need_wgrad = bool(runner._wgrad_enabled)
        ctx.save_for_backward(s0 if has_pre else s1, s1, weights, out, ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        runner, desc = ctx.runner, ctx.desc
        if not desc.training:
            raise RuntimeError("genome_nas_b200: backward through a cell evaluated in eval mode (running-statistics "
                               "BatchNorm) is not implemented; evaluate under torch.no_grad() or call train()")
        s0, s1, weights, out, ws = ctx.saved_tensors
        has_pre = runner._has_pre
        dout = _lib.require(dout.contiguous(), "grad_output")
        region = region_of(runner, device=s1.device)
        gflat = region.ensure_grads() if ctx.need_wgrad else None
        dev = s1.device
        scratch = torch.empty(ctx.bwd_floats, dtype=torch.float32, device=dev)
        ds1 = torch.empty_like(s1)
        ds0 = torch.empty_like(s0) if has_pre else None
        dmix = torch.empty_like(weights)
        _lib.check(lib.gnas_cell_backward(C.byref(desc), _lib.ptr(region.flat), _lib.ptr(gflat),
                                          _lib.ptr(s0) if has_pre else None, _lib.ptr(s1), _lib.ptr(weights),
                                          ctx.mask_arr, _lib.ptr(out), _lib.ptr(dout), _lib.ptr(ws),
                                          _lib.ptr(scratch), _lib.ptr(ds0), _lib.ptr(ds1), _lib.ptr(dmix),
                                          1 if ctx.need_wgrad else 0, _lib.stream_ptr(s1)), "cell_backward")
        _lib.LAUNCHES += 1
        return None, ds0, ds1, dmix, None


class _FusedBase(nn.Module):
    _has_pre = False
    _wgrad_enabled = True     # the first-order Architect switches this off for the alpha pass

    def _anchor(self):
        return next(self.parameters(), None)

    def _device(self):
        t = next(self.parameters(), None)
        return t.device if t is not None else next(self.buffers()).device

    def _nbt(self):
        cache = self.__dict__.get("_gnas_nbt")
        if cache is None or (cache and cache[0].device != self._device()):
            cache = [b for n, b in self.named_buffers() if n.endswith("num_batches_tracked")]
            self.__dict__["_gnas_nbt"] = cache
        return cache

    def _bump_batches_tracked(self):
        nbt = self._nbt()
        if nbt:
            torch._foreach_add_(nbt, 1)

    def _draw_skip_masks(self, desc, like):
        """nn.Dropout(p) behind the identity skips (P-DARTS; model_discCNN.py:60-61).  Like the reference, an edge
        has skip dropout only if its MixedOp was BUILT with p > 0 (Identity wrapped in nn.Sequential(Identity,
        Dropout)); `update_p()` changes the rate of those wrappers and nothing else.  Masks are drawn with torch's
        generator for `like.device`, pre-scaled by 1/(1-p), and applied inside the kernels."""
        if not desc.training:
            return None, None
        masks, arr = [], (C.c_int64 * _lib.MAX_EDGES)()
        for e in range(desc.n_edges):
            arr[e] = 0
            p = self._skip_dropout_p(e)
            if p > 0 and desc.sw[e][3] and self._edge_stride(e) == 1:
                m = torch.empty(desc.B, desc.C, desc.L_out, dtype=torch.float32, device=like.device)
                m.bernoulli_(1.0 - p).div_(1.0 - p)
                masks.append(m)
                arr[e] = m.data_ptr()
        return (masks, arr) if masks else (None, None)


def _skip_p_of(mixed):
    """rate of the Dropout behind the identity skip of one MixedOp, 0 when the wrapper does not exist"""
    op = getattr(mixed.m_ops, 'skip_connect', None)
    if isinstance(op, nn.Sequential) and isinstance(op[0], Identity) and isinstance(op[1], nn.Dropout):
        return float(op[1].p)
    return 0.0


class MixedOp(_FusedBase):
    """sum_k weights[k] * op_k(x) over the enabled primitives (model_discCNN.py:46-93)."""

    def __init__(self, C, stride, switch, p):
        super().__init__()
        self.m_ops = nn.ModuleList()
        self.p = p
        self._C, self._stride, self._switch = C, stride, list(switch)
        for i, on in enumerate(switch):
            if not on:
                continue
            primitive = PRIMITIVES_cnn[i]
            op = OPS[primitive](C, stride, False)
            if 'pool' in primitive:
                op = nn.Sequential(op, nn.BatchNorm1d(C, affine=False))
            if isinstance(op, Identity) and p > 0:
                op = nn.Sequential(op, nn.Dropout(self.p))
            self.m_ops.add_module(primitive, op)

    def update_p(self):
        for op in self.m_ops:
            if isinstance(op, nn.Sequential) and isinstance(op[0], Identity):
                op[1].p = self.p

    def _edge_stride(self, e):
        return self._stride

    def _skip_dropout_p(self, e):
        return _skip_p_of(self)

    def _descriptor(self, region, s0, s1, weights):
        B, Cc, L = s1.shape
        if Cc != self._C:
            raise RuntimeError(f"MixedOp: expected {self._C} channels, got {Cc}")
        d = _lib.CellDesc()
        d.B, d.C, d.C_pp, d.C_p, d.L_pp, d.L_p = B, Cc, Cc, Cc, L, L
        d.stride = self._stride
        d.L_out = (L - 1) // self._stride + 1
        d.n_edges, d.has_pre, d.pre0_factorized = 1, 0, 0
        d.n_prims = sum(self._switch)
        if weights.numel() != d.n_prims:
            raise RuntimeError(f"MixedOp: expected {d.n_prims} mixing weights, got {weights.numel()}")
        d.training = 1 if self.training else 0
        d.bn_momentum = 0.1
        _fill_edge(d, region, 0, "", self._switch, self._stride)
        return d

    def forward(self, x, weights):
        return _FusedCellFunction.apply(self, x, x, weights, self._anchor())


class CNN_Cell_search(_FusedBase):
    """Search cell: preprocess0/1, 4 nodes fed by 2+3+4+5 MixedOp edges, channel concat
    (model_discCNN.py:96-201).  The whole cell is one autograd node backed by gnas_cell_forward/backward."""

    _has_pre = True

    def __init__(self, steps, multiplier, C_prev_prev, C_prev, C, reduction, reduction_prev, switches, p,
                 reduction_high):
        super().__init__()
        if steps != 4 or multiplier != 4:
            raise ValueError("the fused cell kernels implement the reference's steps=4, multiplier=4 cell")
        self.reduction = reduction
        self.p = p
        if reduction_prev:
            self.preprocess0 = FactorizedReduce(C_prev_prev, C, 2, affine=False)
        else:
            self.preprocess0 = ReLUConvBN(C_prev_prev, C, 1, 1, 0, affine=False)
        self.preprocess1 = ReLUConvBN(C_prev, C, 1, 1, 0, affine=False)
        self._steps, self._multiplier = steps, multiplier
        self._C, self._C_pp, self._C_p = C, C_prev_prev, C_prev
        self._pre0_fr = bool(reduction_prev)
        self._cell_stride = 3 if reduction_high else (2 if reduction else 1)
        self._switches = [list(s) for s in switches]
        if len(self._switches) != 14 or any(not any(s) for s in self._switches):
            raise ValueError("fused cell: 14 edges, none of them fully disabled (DARTS / P-DARTS / CWP-DARTS)")
        self.cell_ops = nn.ModuleList()
        e = 0
        for i in range(steps):
            for j in range(2 + i):
                stride = self._cell_stride if j < 2 else 1
                self.cell_ops.add_module(str(e), MixedOp(C, stride, switch=self._switches[e], p=self.p))
                e += 1

    def update_p(self):
        for op in self.cell_ops:
            op.p = self.p
            op.update_p()

    def _edge_stride(self, e):
        return self.cell_ops[e]._stride

    def _skip_dropout_p(self, e):
        return _skip_p_of(self.cell_ops[e])

    def _descriptor(self, region, s0, s1, weights):
        B, c_pp, L_pp = s0.shape
        B1, c_p, L_p = s1.shape
        if (c_pp, c_p) != (self._C_pp, self._C_p) or B != B1:
            raise RuntimeError(f"cell: unexpected input shapes {tuple(s0.shape)} / {tuple(s1.shape)}")
        d = _lib.CellDesc()
        d.B, d.C, d.C_pp, d.C_p, d.L_pp, d.L_p = B, self._C, c_pp, c_p, L_pp, L_p
        d.stride = self._cell_stride
        d.L_out = (L_p - 1) // d.stride + 1
        d.n_edges, d.has_pre, d.pre0_factorized = 14, 1, 1 if self._pre0_fr else 0
        d.n_prims = sum(self._switches[0])
        if tuple(weights.shape) != (14, d.n_prims):
            raise RuntimeError(f"cell: mixing weights must be [14, {d.n_prims}], got {tuple(weights.shape)}")
        d.training = 1 if self.training else 0
        d.bn_momentum = 0.1
        for e in range(14):
            _fill_edge(d, region, e, f"cell_ops.{e}.", self._switches[e], self.cell_ops[e]._stride)
        if self._pre0_fr:
            d.pre_w_off[0] = region.offsets['preprocess0.conv_1.weight']
            d.pre_bn_off[0] = region.bn_offsets['preprocess0.bn.running_mean']
        else:
            d.pre_w_off[0] = region.offsets['preprocess0.op.1.weight']
            d.pre_bn_off[0] = region.bn_offsets['preprocess0.op.2.running_mean']
        d.pre_w_off[1] = region.offsets['preprocess1.op.1.weight']
        d.pre_bn_off[1] = region.bn_offsets['preprocess1.op.2.running_mean']
        return d

    def forward(self, s0, s1, weights):
        return _FusedCellFunction.apply(self, s0, s1, weights, self._anchor())
