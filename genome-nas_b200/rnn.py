"""DARTSCellSearch -- the recurrent highway layer of the supernet -- with the reference's module API,
evaluated by the persistent sm_100a kernel (csrc/rhn.cu) through the C ABI.

Reference interface mirrored: darts_tools/model_discCNN.py:209-268 (DARTSCell.__init__/forward,
_compute_init_state) and model_search.py:36-97 (DARTSCellSearch): parameters `_W0`, `_Ws.0..7`,
module `bn`, attribute `weights` assigned from outside, `forward(inputs[T,B,ninp], hidden[1,B,nhid])
-> (hiddens[T,B,nhid], hiddens[-1][None])`.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .flat import region_of
from .operations import PRIMITIVES_rnn, rnn_steps

INITRANGE = 0.04


def mask2d(B, D, keep_prob, device):
    """Variational-dropout mask, drawn on the CPU generator exactly like generalNAS_tools/utils.py:354-363
    (quirk Q10) so that seeded runs consume the same random stream as the reference."""
    m = torch.floor(torch.rand(B, D) + keep_prob) / keep_prob
    return m.to(device)


class _RhnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, h0, probs, x_mask, h_mask, _anchor):
        lib = _lib.lib()
        region = region_of(mod, exclude=('weights',))
        x = _lib.require(x.contiguous(), "inputs")
        h0 = _lib.require(h0.contiguous(), "hidden")
        probs = _lib.require(probs.detach().contiguous(), "probs")
        T, B, H = x.shape
        desc = mod._descriptor(T, B, H, probs)
        fwd, bwd = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.gnas_rhn_workspace(C.byref(desc), C.byref(fwd), C.byref(bwd)), "rhn_workspace")
        ws = torch.empty(fwd.value, dtype=torch.float32, device=x.device)
        out = torch.empty(T, B, H, dtype=torch.float32, device=x.device)
        flat, off = region.flat, region.offsets
        w0 = flat[off["_W0"]:]
        wsx = flat[off["_Ws.0"]:]
        if off["_Ws.7"] - off["_Ws.0"] != 7 * H * 2 * H:
            raise RuntimeError("_Ws.0..7 must be contiguous in the flat buffer")
        bn = region.bn[region.bn_offsets["bn.running_mean"]:]
        if x_mask is not None:
            x_mask = _lib.require(x_mask.contiguous(), "x_mask")
            h_mask = _lib.require(h_mask.contiguous(), "h_mask")
        _lib.check(lib.gnas_rhn_forward(C.byref(desc), _lib.ptr(x), _lib.ptr(h0), _lib.ptr(w0), _lib.ptr(wsx),
                                        _lib.ptr(probs), _lib.ptr(x_mask), _lib.ptr(h_mask), _lib.ptr(bn),
                                        _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr(x)), "rhn_forward")
        _lib.LAUNCHES += 1
        if desc.training:
            mod.bn.num_batches_tracked += 9 * T
        ctx.mod, ctx.desc, ctx.bwd_floats = mod, desc, bwd.value
        ctx.need_wgrad = bool(mod._wgrad_enabled)
        ctx.masks = (x_mask, h_mask)
        ctx.save_for_backward(x, h0, probs, ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        mod, desc = ctx.mod, ctx.desc
        if not desc.training:
            raise RuntimeError("genome_nas_b200: backward through the RHN layer evaluated in eval mode (running-"
                               "statistics BatchNorm) is not implemented; evaluate under torch.no_grad() or call train()")
        x, h0, probs, ws = ctx.saved_tensors
        x_mask, h_mask = ctx.masks
        dout = _lib.require(dout.contiguous(), "grad_output")
        region = region_of(mod, exclude=('weights',))
        flat, off = region.flat, region.offsets
        dw0 = dws = None
        if ctx.need_wgrad:
            gflat = region.ensure_grads()
            dw0, dws = gflat[off["_W0"]:], gflat[off["_Ws.0"]:]
        scratch = torch.empty(ctx.bwd_floats, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        dh0 = torch.empty_like(h0)
        dprobs = torch.empty_like(probs)
        _lib.check(lib.gnas_rhn_backward(C.byref(desc), _lib.ptr(x), _lib.ptr(h0), _lib.ptr(flat[off["_W0"]:]),
                                         _lib.ptr(flat[off["_Ws.0"]:]), _lib.ptr(probs), _lib.ptr(x_mask),
                                         _lib.ptr(h_mask), _lib.ptr(dout), _lib.ptr(ws), _lib.ptr(scratch),
                                         _lib.ptr(dx), _lib.ptr(dh0), _lib.ptr(dw0), _lib.ptr(dws), _lib.ptr(dprobs),
                                         1 if ctx.need_wgrad else 0, _lib.stream_ptr(x)), "rhn_backward")
        _lib.LAUNCHES += 1
        return None, dx, dh0, dprobs, None, None, None


class DARTSCellSearch(nn.Module):
    """Search-time RHN cell: mixed tanh / identity / sigmoid / relu highway edges, one shared
    BatchNorm1d(nhid, affine=False) for all nine states (model_search.py:36-97)."""

    _wgrad_enabled = True

    def __init__(self, ninp, nhid, dropouth, dropoutx, switch_rnn):
        super().__init__()
        if ninp != nhid:
            raise ValueError("the supernet uses ninp == nhid (model_discCNN.py:421)")
        self.nhid, self.dropouth, self.dropoutx = nhid, dropouth, dropoutx
        self.genotype = None
        # same construction order and RNG calls as DARTSCell.__init__ (model_discCNN.py:219-223)
        self._W0 = nn.Parameter(torch.Tensor(ninp + nhid, 2 * nhid).uniform_(-INITRANGE, INITRANGE))
        self._Ws = nn.ParameterList(
            [nn.Parameter(torch.Tensor(nhid, 2 * nhid).uniform_(-INITRANGE, INITRANGE)) for _ in range(rnn_steps)])
        self.bn = nn.BatchNorm1d(nhid, affine=False)
        self.switch_rnn = [list(s) for s in switch_rnn]
        if len(self.switch_rnn) != 36 or any(len(s) != len(PRIMITIVES_rnn) for s in self.switch_rnn):
            raise ValueError("switch_rnn must be 36 x 5")

    def _descriptor(self, T, B, H, probs):
        if H != self.nhid:
            raise RuntimeError(f"RHN: expected hidden size {self.nhid}, got {H}")
        d = _lib.RhnDesc()
        d.T, d.B, d.H = T, B, H
        d.training = 1 if self.training else 0
        d.n_prims = probs.shape[1]
        d.bn_momentum = 0.1
        for r, row in enumerate(self.switch_rnn):
            for k, on in enumerate(row):
                d.sw[r][k] = 1 if on else 0
        # the backward kernel writes 36 * n_prims gradient entries: the matrix must be exactly [36, n_prims]
        # (DARTS / P-DARTS / CWP-DARTS never disable a whole RNN edge)
        if tuple(probs.shape) != (36, d.n_prims) or any(not any(row) or sum(row) > d.n_prims for row in self.switch_rnn):
            raise RuntimeError(f"RHN: probability matrix must be [36, n_prims] and match switch_rnn, got {tuple(probs.shape)}")
        return d

    def forward(self, inputs, hidden):
        T, B = inputs.size(0), inputs.size(1)
        provider = self.__dict__.get("_mask_provider")   # SearchStep's CUDA-graph mode feeds static mask buffers
        if self.training and provider is not None:
            x_mask, h_mask = provider(B, inputs.size(2))
        elif self.training:
            x_mask = mask2d(B, inputs.size(2), 1. - self.dropoutx, inputs.device)
            h_mask = mask2d(B, hidden.size(2), 1. - self.dropouth, inputs.device)
        else:
            x_mask = h_mask = None
        probs = F.softmax(self.weights, dim=-1)
        hiddens = _RhnFunction.apply(self, inputs, hidden[0], probs, x_mask, h_mask, self._W0)
        return hiddens, hiddens[-1].unsqueeze(0)
