"""RNNModel / RNNModelSearch: the genomeDARTS supernet behind the reference's module API.

Reference interface mirrored: darts_tools/model_discCNN.py:307-581 (RNNModel) and
model_search.py:113-243 (RNNModelSearch): the 20 positional constructor arguments, `forward(input,
hidden, return_h)`, `_loss`, `new`, `arch_parameters`, `genotype`, `init_hidden`, `update_p`, the
attributes `alphas_normal / alphas_reduce / weights / p`, and every state_dict key (3,933 entries
for the full supernet), with parameters created in the reference's order so that a seeded
construction yields the same initial weights.

Stem, the six search cells and the recurrent highway layer run on the sm_100a kernels (C ABI,
include/gnas.h).  The head (Dropout -> Linear(21504,925) -> ReLU -> Linear(925,919) -> Sigmoid) is
plain library GEMM work and stays on torch/cuBLAS in fp32 with TF32 disabled.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .cnn import CNN_Cell_search
from .flat import region_of
from .head import criterion_loss, head_forward
from .operations import CONCAT, PRIMITIVES_cnn, PRIMITIVES_rnn, rnn_steps
from .rnn import INITRANGE, DARTSCellSearch
from collections import namedtuple

Genotype = namedtuple('Genotype', 'normal normal_concat reduce reduce_concat rnn rnn_concat')   # genotypes.py:11


def encode_onehot(x):
    """Host-side: fp32 one-hot [B, 4, L] -> uint8 base codes [B, L] (0..3 = hot channel, 255 = all-zero column)."""
    hot = x.max(dim=1)
    return torch.where(hot.values > 0, hot.indices, torch.full_like(hot.indices, 255)).to(torch.uint8).contiguous()


def expand_onehot(codes):
    """Device-side inverse of `encode_onehot`: uint8 codes [B, L] -> fp32 one-hot [B, 4, L] (gnas_onehot_expand)."""
    codes = _lib.require(codes.contiguous(), "base codes", torch.uint8)
    B, L = codes.shape
    x = torch.empty(B, 4, L, dtype=torch.float32, device=codes.device)
    _lib.check(_lib.lib().gnas_onehot_expand(_lib.ptr(codes), B, L, _lib.ptr(x), _lib.stream_ptr(codes)), "onehot_expand")
    _lib.LAUNCHES += 1
    return x


class _StemFunction(torch.autograd.Function):
    """Conv1d(4, C0, 13, pad 6) + BatchNorm1d(C0) affine (model_discCNN.py:350-353)."""

    @staticmethod
    def forward(ctx, stem, x, _anchor):
        lib = _lib.lib()
        region = region_of(stem)
        x = _lib.require(x.contiguous(), "input")
        B, Cin, L = x.shape
        if Cin != 4:
            raise RuntimeError("stem expects a one-hot input with 4 channels")
        conv, bn = stem[0], stem[1]
        C0 = conv.out_channels
        fwd, bwd = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.gnas_stem_workspace(B, C0, L, C.byref(fwd), C.byref(bwd)), "stem_workspace")
        ws = torch.empty(fwd.value, dtype=torch.float32, device=x.device)
        out = torch.empty(B, C0, L, dtype=torch.float32, device=x.device)
        training = 1 if bn.training else 0
        _lib.check(lib.gnas_stem_forward(B, C0, L, training, 0.1, _lib.ptr(x), _lib.ptr(conv.weight),
                                         _lib.ptr(bn.weight), _lib.ptr(bn.bias), _lib.ptr(region.bn), _lib.ptr(out),
                                         _lib.ptr(ws), _lib.stream_ptr(x)), "stem_forward")
        _lib.LAUNCHES += 1
        if training:
            bn.num_batches_tracked += 1
        ctx.stem, ctx.dims, ctx.bwd_floats, ctx.training = stem, (B, C0, L), bwd.value, training
        ctx.need_wgrad = bool(getattr(stem, "_wgrad_enabled", True))
        ctx.save_for_backward(x, ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        if not ctx.need_wgrad:
            return None, None, None
        if not ctx.training:
            raise RuntimeError("genome_nas_b200: backward through the stem evaluated in eval mode (running-statistics "
                               "BatchNorm) is not implemented; evaluate under torch.no_grad() or call train()")
        lib = _lib.lib()
        stem = ctx.stem
        x, ws = ctx.saved_tensors
        B, C0, L = ctx.dims
        region = region_of(stem)
        g = region.ensure_grads()
        off = region.offsets
        dout = _lib.require(dout.contiguous(), "grad_output")
        scratch = torch.empty(ctx.bwd_floats, dtype=torch.float32, device=x.device)
        _lib.check(lib.gnas_stem_backward(B, C0, L, _lib.ptr(x), _lib.ptr(stem[0].weight), _lib.ptr(stem[1].weight),
                                          _lib.ptr(dout), _lib.ptr(ws), _lib.ptr(scratch), _lib.ptr(g[off["0.weight"]:]),
                                          _lib.ptr(g[off["1.weight"]:]), _lib.ptr(g[off["1.bias"]:]),
                                          _lib.stream_ptr(x)), "stem_backward")
        _lib.LAUNCHES += 1
        return None, None, None


class RNNModel(nn.Module):

    def __init__(self, seq_len, dropouth=0.5, dropoutx=0.5, C=8, num_classes=4, layers=3, steps=4, multiplier=4,
                 stem_multiplier=3, search=True, drop_path_prob=0.2, genotype=None, task=None, switches_normal=[],
                 switches_reduce=[], switches_rnn=[], p=0.0, alphas_normal=None, alphas_reduce=None, alphas_rnn=None):
        super().__init__()
        if not search or genotype is not None:
            raise NotImplementedError("genome_nas_b200 implements the search supernet (search=True, genotype=None); "
                                      "derived-architecture training is outside the accelerated path")
        if layers != 6:
            raise NotImplementedError("the reference hard-codes the stride-3 cell at index 5: layers must be 6")
        self._C, self._num_classes, self._layers = C, num_classes, layers
        self._steps, self._multiplier = steps, multiplier
        self.p, self.dropouth, self.dropoutx = p, dropouth, dropoutx
        self.switches_normal, self.switches_reduce, self.switches_rnn = switches_normal, switches_reduce, switches_rnn
        self.num_ops_cnn = sum(switches_normal[0])
        self.num_ops_rnn = sum(switches_rnn[0])

        C_curr = stem_multiplier * C
        self.stem = nn.Sequential(nn.Conv1d(4, C_curr, 13, padding=6, bias=False), nn.BatchNorm1d(C_curr))
        C_pp, C_p, C_curr = C_curr, C_curr, C
        self.cells = nn.ModuleList()
        reduction_prev = False
        neurons = seq_len
        for i in range(layers):
            reduction = i % 3 != 0                 # normal cells at 0 and 3 (model_discCNN.py:365)
            high = reduction and i == 5            # stride-3 reduction (model_discCNN.py:376)
            if reduction:
                C_curr *= 2
                neurons = round(neurons / (3 if high else 2))
            switches = self.switches_reduce if reduction else self.switches_normal
            self.cells.append(CNN_Cell_search(steps, multiplier, C_pp, C_p, C_curr, reduction, reduction_prev,
                                              switches, self.p, high))
            reduction_prev = reduction
            C_pp, C_p = C_p, multiplier * C_curr
        self.drop_path_prob = drop_path_prob
        out_channels = C_curr * steps
        self.rnns = nn.ModuleList([DARTSCellSearch(out_channels, out_channels, dropouth, dropoutx, self.switches_rnn)])
        self.alphas_normal = alphas_normal
        self.alphas_reduce = alphas_reduce
        self.weights = alphas_rnn
        for rnn in self.rnns:
            rnn.weights = self.weights
        self._arch_parameters = [self.alphas_normal, self.alphas_reduce, self.weights]
        self.decoder = nn.Linear(neurons * out_channels, 925)
        self.classifier = nn.Linear(925, num_classes)
        self.final = nn.Sigmoid() if task == "TF_bindings" else nn.Identity()
        self.relu = nn.ReLU(inplace=True)
        self.dropout_lin = nn.Dropout(p=0.1)
        self.init_weights()
        self.ninp = self.nhid = self.nhidlast = out_channels

    _head_wgrad = True

    def init_weights(self):
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.uniform_(-INITRANGE, INITRANGE)

    # -- kernels' view of the model -------------------------------------------------------------
    def flat_region(self):
        """One contiguous fp32 buffer under all parameters (and one under all gradients)."""
        return region_of(self)

    def set_weight_grads(self, enabled: bool):
        """Disable the weight-gradient kernels for passes whose weight gradients are discarded
        (the first-order alpha step: architect.py:69-72 + train_and_validate.py:112)."""
        self.stem._wgrad_enabled = enabled
        self._head_wgrad = enabled
        for cell in self.cells:
            cell._wgrad_enabled = enabled
        for rnn in self.rnns:
            rnn._wgrad_enabled = enabled

    @staticmethod
    def _softmax_alpha(alpha):
        return F.softmax(alpha, dim=0 if alpha.size(1) == 1 else -1)     # model_discCNN.py:497-507

    def forward(self, input, hidden, return_h=False):
        region_of(self)
        if input.dtype == torch.uint8:          # compact data path: base codes [B, L], expanded on the device
            input = expand_onehot(input)
        s0 = s1 = _StemFunction.apply(self.stem, input, self.stem[0].weight)
        w_normal = w_reduce = None
        for cell in self.cells:
            if cell.reduction:
                if w_reduce is None:
                    w_reduce = self._softmax_alpha(self.alphas_reduce)
                weights = w_reduce
            else:
                if w_normal is None:
                    w_normal = self._softmax_alpha(self.alphas_normal)
                weights = w_normal
            s0, s1 = s1, cell(s0, s1, weights)
        raw_output = s1.permute(2, 0, 1)            # [T, B, nhid]
        new_hidden, raw_outputs, outputs = [], [], []
        for l, rnn in enumerate(self.rnns):
            raw_output, new_h = rnn(raw_output, hidden[l])
            new_hidden.append(new_h)
            raw_outputs.append(raw_output)
        hidden = new_hidden
        output = raw_output
        outputs.append(output)
        logit = head_forward(self, output)          # permute/flatten, dropout_lin, decoder, relu, classifier, final
        if return_h:
            return logit, hidden, raw_outputs, outputs
        return logit, hidden

    def update_p(self):
        for cell in self.cells:
            cell.p = self.p
            cell.update_p()

    def init_hidden(self, bsz):
        weight = next(self.parameters()).data
        return [weight.new_zeros(1, bsz, self.nhid)]


class RNNModelSearch(RNNModel):

    def __init__(self, *args):
        super().__init__(*args)
        self._args = args

    def new(self):
        model_new = RNNModelSearch(*self._args)
        for x, y in zip(model_new.arch_parameters(), self.arch_parameters()):
            x.data.copy_(y.data)
        return model_new

    def _loss(self, hidden, input, target, criterion):
        log_prob, hidden_next = self(input, hidden, return_h=False)
        return criterion_loss(criterion, log_prob, target), hidden_next

    def arch_parameters(self):
        return self._arch_parameters

    def genotype(self):
        """Host-side derivation of the discrete architecture (model_search.py:164-243): per node the two
        strongest input edges (CNN) / the strongest predecessor (RNN), ranked by their best non-'none' weight."""
        none_c, none_r = PRIMITIVES_cnn.index('none'), PRIMITIVES_rnn.index('none')

        def best_op(row, none_idx):
            ks = [k for k in range(len(row)) if k != none_idx]
            return max(ks, key=lambda k: (row[k], -k))

        def strength(row, none_idx):
            return max(row[k] for k in range(len(row)) if k != none_idx)

        def parse_cnn(w):
            gene, start = [], 0
            for i in range(self._steps):
                rows = w[start:start + i + 2]
                order = sorted(range(i + 2), key=lambda j: -strength(rows[j], none_c))[:2]
                gene += [(PRIMITIVES_cnn[best_op(rows[j], none_c)], j) for j in order]
                start += i + 2
            return gene

        def parse_rnn(w):
            gene, start = [], 0
            for i in range(rnn_steps):
                rows = w[start:start + i + 1]
                j = sorted(range(i + 1), key=lambda e: -strength(rows[e], none_r))[0]
                gene.append((PRIMITIVES_rnn[best_op(rows[j], none_r)], j))
                start += i + 1
            return gene

        sm = lambda a: F.softmax(a, dim=-1).data.cpu().numpy()
        concat = range(2 + self._steps - self._multiplier, self._steps + 2)
        return Genotype(normal=parse_cnn(sm(self.alphas_normal)), normal_concat=concat,
                        reduce=parse_cnn(sm(self.alphas_reduce)), reduce_concat=concat,
                        rnn=parse_rnn(sm(self.weights)), rnn_concat=range(rnn_steps + 1)[-CONCAT:])
