"""Mask-based supernet of OSP-NAS / Hyperband-NAS / random search on the fused kernels.

Reference interface mirrored: randomSearch_and_Hyperband_Tools/model.py:32-141 (MixedOp, CNN_Cell_search),
:248-471 (RNNModel: `forward(input, hidden, mask, return_h)`, `mask = [normal 14x9, reduce 14x9, rnn 36x5]` numpy
0/1 matrices, `dropout_lin` p = 0.5) and randomSearch_and_Hyperband_Tools/model_search.py:34-86 (DARTSCellSearch
with `forward(inputs, hidden, rnn_mask)`), :93-112 (RNNModelSearch).  Same child-module names (`_ops.<index>` with
None for primitives outside the supernet mask) and hence the same state_dict keys, so OSP-NAS can keep copying
weights between supernets by key (train_genomicOSP_NAS.py:404-410).

The arithmetic is the DARTS hot path with mixing weights in {0, 1} and no architecture gradient: a MixedOp is the
UNWEIGHTED sum of the candidates the call-time mask selects, an RHN edge adds s_j + c (f(h) - s_j) once per selected
activation.  Only the selected candidates are evaluated (their BatchNorm running statistics alone are updated,
like in the reference).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .cnn import _FusedBase, _FusedCellFunction, _fill_edge
from .model import _StemFunction
from .flat import region_of
from .head import criterion_loss, head_forward
from .operations import OPS, PRIMITIVES_cnn, PRIMITIVES_rnn, FactorizedReduce, ReLUConvBN, rnn_steps
from .rnn import INITRANGE, _RhnFunction, mask2d  # noqa: F401  (mask2d: same CPU-generator draw order)
from . import rnn as _rnn


def _as_mask(m, rows, cols):
    a = np.asarray(m)
    if a.shape != (rows, cols):
        raise RuntimeError(f"mask must be {rows} x {cols}, got {a.shape}")
    return a != 0


class MixedOp(_FusedBase):
    """model.py:32-82: `_ops` holds one child per primitive index, None outside the supernet mask."""

    def __init__(self, C, stride, mask_cnn):
        super().__init__()
        self.stride = stride
        self._C = C
        self._ops = nn.ModuleList()
        self._super = [bool(v) for v in np.asarray(mask_cnn) != 0]
        for i, on in enumerate(self._super):
            op = None
            if on:
                primitive = PRIMITIVES_cnn[i]
                op = OPS[primitive](C, stride, False)
                if 'pool' in primitive:
                    op = nn.Sequential(op, nn.BatchNorm1d(C, affine=False))
            self._ops.add_module(str(i), op)
        self._call = None

    p = 0.0

    def _edge_stride(self, e):
        return self.stride

    def _skip_dropout_p(self, e):
        return 0.0

    def _descriptor(self, region, s0, s1, weights):
        B, Cc, L = s1.shape
        sel = [a and b for a, b in zip(self._super, self._call)]
        d = _lib.CellDesc()
        d.B, d.C, d.C_pp, d.C_p, d.L_pp, d.L_p = B, Cc, Cc, Cc, L, L
        d.stride = self.stride
        d.L_out = (L - 1) // self.stride + 1
        d.n_edges, d.has_pre, d.pre0_factorized = 1, 0, 0
        d.n_prims = len(PRIMITIVES_cnn)
        d.training = 1 if self.training else 0
        d.bn_momentum = 0.1
        _fill_edge(d, region, 0, "", sel, self.stride, holder="_ops", by_index=True)
        return d

    def forward(self, x, mask_cnn):
        call = np.asarray(mask_cnn) != 0
        if any(c and not s for c, s in zip(call, self._super)):
            raise RuntimeError("MixedOp: the call-time mask selects a primitive outside the supernet mask")
        self._call = [bool(v) for v in call]
        ones = torch.ones(len(PRIMITIVES_cnn), dtype=torch.float32, device=x.device)
        return _FusedCellFunction.apply(self, x, x, ones, self._anchor())


class CNN_Cell_search(_FusedBase):
    """model.py:85-141."""

    _has_pre = True
    p = 0.0

    def __init__(self, steps, multiplier, C_prev_prev, C_prev, C, reduction, reduction_prev, mask_cnn, reduction_high):
        super().__init__()
        if steps != 4 or multiplier != 4:
            raise ValueError("the fused cell kernels implement the reference's steps=4, multiplier=4 cell")
        self.reduction = reduction
        if reduction_prev:
            self.preprocess0 = FactorizedReduce(C_prev_prev, C, 2, affine=False)
        else:
            self.preprocess0 = ReLUConvBN(C_prev_prev, C, 1, 1, 0, affine=False)
        self.preprocess1 = ReLUConvBN(C_prev, C, 1, 1, 0, affine=False)
        self._steps, self._multiplier = steps, multiplier
        self._C, self._C_pp, self._C_p = C, C_prev_prev, C_prev
        self._pre0_fr = bool(reduction_prev)
        self._cell_stride = 3 if reduction_high else (2 if reduction else 1)
        mask = _as_mask(mask_cnn, 14, len(PRIMITIVES_cnn))
        self.cell_ops = nn.ModuleList()
        self._bns = nn.ModuleList()
        e = 0
        for i in range(steps):
            for j in range(2 + i):
                self.cell_ops.add_module(str(e), MixedOp(C, self._cell_stride if j < 2 else 1, mask[e]))
                e += 1
        self._call = None

    def _edge_stride(self, e):
        return self.cell_ops[e].stride

    def _skip_dropout_p(self, e):
        return 0.0

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
        d.n_prims = len(PRIMITIVES_cnn)
        d.training = 1 if self.training else 0
        d.bn_momentum = 0.1
        for e in range(14):
            sup = self.cell_ops[e]._super
            sel = [bool(a and b) for a, b in zip(sup, self._call[e])]
            _fill_edge(d, region, e, f"cell_ops.{e}.", sel, self.cell_ops[e].stride, holder="_ops", by_index=True)
        if self._pre0_fr:
            d.pre_w_off[0] = region.offsets['preprocess0.conv_1.weight']
            d.pre_bn_off[0] = region.bn_offsets['preprocess0.bn.running_mean']
        else:
            d.pre_w_off[0] = region.offsets['preprocess0.op.1.weight']
            d.pre_bn_off[0] = region.bn_offsets['preprocess0.op.2.running_mean']
        d.pre_w_off[1] = region.offsets['preprocess1.op.1.weight']
        d.pre_bn_off[1] = region.bn_offsets['preprocess1.op.2.running_mean']
        return d

    def _nbt(self):
        """num_batches_tracked of the BatchNorms that actually ran: preprocess + the selected candidates"""
        out = [b for n, b in self.named_buffers() if n.endswith("num_batches_tracked") and n.startswith("preprocess")]
        for e in range(14):
            for k, on in enumerate(self._call[e]):
                if on and self.cell_ops[e]._super[k]:
                    out += [b for n, b in self.cell_ops[e]._ops[k].named_buffers() if n.endswith("num_batches_tracked")]
        return out

    def forward(self, s0, s1, mask_cnn):
        call = _as_mask(mask_cnn, 14, len(PRIMITIVES_cnn))
        for e in range(14):
            if any(c and not s for c, s in zip(call[e], self.cell_ops[e]._super)):
                raise RuntimeError(f"cell: the call-time mask of edge {e} selects a primitive outside the supernet mask")
        self._call = call
        ones = torch.ones(14, len(PRIMITIVES_cnn), dtype=torch.float32, device=s1.device)
        return _FusedCellFunction.apply(self, s0, s1, ones, self._anchor())


class DARTSCellSearch(_rnn.DARTSCellSearch):
    """model_search.py:34-86: every (edge, activation) the 36 x 5 mask selects contributes with weight one."""

    def __init__(self, ninp, nhid, dropouth, dropoutx):
        super().__init__(ninp, nhid, dropouth, dropoutx, [[True] * len(PRIMITIVES_rnn) for _ in range(36)])
        self.weights = None

    def _descriptor(self, T, B, H, probs):
        if H != self.nhid:
            raise RuntimeError(f"RHN: expected hidden size {self.nhid}, got {H}")
        d = _lib.RhnDesc()
        d.T, d.B, d.H = T, B, H
        d.training = 1 if self.training else 0
        d.n_prims = len(PRIMITIVES_rnn)
        d.bn_momentum = 0.1
        for r in range(36):
            for k in range(len(PRIMITIVES_rnn)):
                d.sw[r][k] = 1 if self._call[r][k] else 0
        return d

    def forward(self, inputs, hidden, rnn_mask):
        call = _as_mask(rnn_mask, 36, len(PRIMITIVES_rnn))
        if call[:, 0].any():
            raise NotImplementedError("the mask selects 'none' as an RHN activation (the reference raises too: "
                                      "model.py:_get_activation)")
        self._call = call
        T, B = inputs.size(0), inputs.size(1)
        if self.training:
            x_mask = _rnn.mask2d(B, inputs.size(2), 1. - self.dropoutx, inputs.device)
            h_mask = _rnn.mask2d(B, hidden.size(2), 1. - self.dropouth, inputs.device)
        else:
            x_mask = h_mask = None
        ones = torch.ones(36, len(PRIMITIVES_rnn), dtype=torch.float32, device=inputs.device)
        hiddens = _RhnFunction.apply(self, inputs, hidden[0], ones, x_mask, h_mask, self._W0)
        return hiddens, hiddens[-1].unsqueeze(0)


class RNNModel(nn.Module):
    _head_wgrad = True

    """model.py:248-471 (search=True)."""

    def __init__(self, seq_len, dropouth=0.5, dropoutx=0.5, C=8, num_classes=4, layers=3, steps=4, multiplier=4,
                 stem_multiplier=3, search=True, drop_path_prob=0.2, genotype=None, task=None, mask=None):
        super().__init__()
        if not search or genotype is not None:
            raise NotImplementedError("mask supernet: search=True only (derived architectures: genome_nas_b200.derived)")
        if layers != 6:
            raise NotImplementedError("the reference hard-codes the stride-3 cell at index 5: layers must be 6")
        self.mask_normal, self.mask_reduce, self.mask_rnn = mask[0], mask[1], mask[2]
        self._C, self._num_classes, self._layers, self._steps, self._multiplier = C, num_classes, layers, steps, multiplier
        self.dropouth, self.dropoutx = dropouth, dropoutx
        C_curr = stem_multiplier * C
        self.stem = nn.Sequential(nn.Conv1d(4, C_curr, 13, padding=6, bias=False), nn.BatchNorm1d(C_curr))
        C_pp, C_p, C_curr = C_curr, C_curr, C
        self.cells = nn.ModuleList()
        reduction_prev = False
        neurons = seq_len
        for i in range(layers):
            reduction = i % 3 != 0
            high = reduction and i == 5
            if reduction:
                C_curr *= 2
                neurons = round(neurons / (3 if high else 2))
            self.cells.append(CNN_Cell_search(steps, multiplier, C_pp, C_p, C_curr, reduction, reduction_prev,
                                              self.mask_reduce if reduction else self.mask_normal, high))
            reduction_prev = reduction
            C_pp, C_p = C_p, multiplier * C_curr
        self.drop_path_prob = drop_path_prob
        out_channels = C_curr * steps
        self.rnns = nn.ModuleList([DARTSCellSearch(out_channels, out_channels, dropouth, dropoutx)])
        self.decoder = nn.Linear(neurons * out_channels, 925)
        self.classifier = nn.Linear(925, num_classes)
        self.relu = nn.ReLU(inplace=True)
        self.dropout_lin = nn.Dropout(p=0.5)
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.uniform_(-INITRANGE, INITRANGE)
        self.ninp = self.nhid = self.nhidlast = out_channels
        self.final = nn.Sigmoid() if task == "TF_bindings" else nn.Identity()

    def forward(self, input, hidden, mask, return_h=False):
        region_of(self)
        s0 = s1 = _StemFunction.apply(self.stem, input, self.stem[0].weight)
        for cell in self.cells:
            s0, s1 = s1, cell(s0, s1, mask[1] if cell.reduction else mask[0])
        raw_output = s1.permute(2, 0, 1)
        new_hidden, raw_outputs, outputs = [], [], []
        for l, rnn in enumerate(self.rnns):
            raw_output, new_h = rnn(raw_output, hidden[l], mask[2])
            new_hidden.append(new_h)
            raw_outputs.append(raw_output)
        hidden = new_hidden
        output = raw_output
        outputs.append(output)
        logit = head_forward(self, output)
        if return_h:
            return logit, hidden, raw_outputs, outputs
        return logit, hidden

    def init_hidden(self, bsz):
        weight = next(self.parameters()).data
        return [weight.new_zeros(1, bsz, self.nhid)]


class RNNModelSearch(RNNModel):
    """model_search.py:93-112."""

    def __init__(self, *args):
        super().__init__(*args)
        self._args = args

    def new(self):
        return RNNModelSearch(*self._args)

    def _loss(self, hidden, input, target, criterion, mask=None):
        log_prob, hidden_next = self(input, hidden, mask if mask is not None else
                                     [self.mask_normal, self.mask_reduce, self.mask_rnn], return_h=False)
        return criterion_loss(criterion, log_prob, target), hidden_next
