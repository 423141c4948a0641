"""Classification head and loss terms of the search step on the hand-written kernels (csrc/head.cu).

Reference: darts_tools/model_discCNN.py:547-566 (permute, flatten, Dropout(0.1), Linear(T*nhid, 925), ReLU,
Linear(925, classes), Sigmoid) and generalNAS_tools/train_and_validate.py:136-140 (BCELoss + beta * TAR).
The modules stay `nn.Linear` / `nn.Dropout` (state_dict contract); only their arithmetic moves.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from .flat import region_of


class _HeadFunction(torch.autograd.Function):
    """logit = final(classifier(relu(decoder(dropout(flatten(hs.permute(1, 0, 2)))))));  weight gradients are
    accumulated straight into the flat gradient buffer (like every other fused module of the supernet)."""

    @staticmethod
    def forward(ctx, model, hs, drop_mask, _anchor):
        lib = _lib.lib()
        region = region_of(model)
        hs = _lib.require(hs.contiguous(), "rnn output")
        T, B, H = hs.shape
        dec, cls = model.decoder, model.classifier
        D1, NC = dec.out_features, cls.out_features
        if dec.in_features != T * H or cls.in_features != D1:
            raise RuntimeError(f"head: decoder expects {dec.in_features} features, the RNN output has {T}x{H}")
        sig = 1 if isinstance(model.final, torch.nn.Sigmoid) else 0
        fwd, bwd = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.gnas_head_workspace(T, B, H, D1, NC, C.byref(fwd), C.byref(bwd)), "head_workspace")
        ws = torch.empty(fwd.value, dtype=torch.float32, device=hs.device)
        out = torch.empty(B, NC, dtype=torch.float32, device=hs.device)
        if drop_mask is not None:
            drop_mask = _lib.require(drop_mask.contiguous(), "dropout mask")
        flat, off = region.flat, region.offsets
        wd, bd = flat[off["decoder.weight"]:], flat[off["decoder.bias"]:] if dec.bias is not None else None
        wc, bc = flat[off["classifier.weight"]:], flat[off["classifier.bias"]:] if cls.bias is not None else None
        _lib.check(lib.gnas_head_forward(T, B, H, D1, NC, sig, _lib.ptr(hs), _lib.ptr(drop_mask), _lib.ptr(wd),
                                         _lib.ptr(bd), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(out), _lib.ptr(ws),
                                         _lib.stream_ptr(hs)), "head_forward")
        _lib.LAUNCHES += 1
        ctx.model, ctx.dims, ctx.bwd_floats = model, (T, B, H, D1, NC, sig), bwd.value
        ctx.need_wgrad = bool(model._head_wgrad)
        ctx.drop_mask = drop_mask
        ctx.save_for_backward(out, ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        model = ctx.model
        out, ws = ctx.saved_tensors
        T, B, H, D1, NC, sig = ctx.dims
        dout = _lib.require(dout.contiguous(), "grad_output")
        region = region_of(model)
        flat, off = region.flat, region.offsets
        dwd = dbd = dwc = dbc = None
        if ctx.need_wgrad:
            g = region.ensure_grads()
            dwd, dwc = g[off["decoder.weight"]:], g[off["classifier.weight"]:]
            dbd = g[off["decoder.bias"]:] if model.decoder.bias is not None else None
            dbc = g[off["classifier.bias"]:] if model.classifier.bias is not None else None
        scratch = torch.empty(ctx.bwd_floats, dtype=torch.float32, device=out.device)
        dhs = torch.empty(T, B, H, dtype=torch.float32, device=out.device)
        _lib.check(lib.gnas_head_backward(T, B, H, D1, NC, sig, _lib.ptr(ctx.drop_mask),
                                          _lib.ptr(flat[off["decoder.weight"]:]), _lib.ptr(flat[off["classifier.weight"]:]),
                                          _lib.ptr(out), _lib.ptr(dout), _lib.ptr(ws), _lib.ptr(scratch), _lib.ptr(dhs),
                                          _lib.ptr(dwd), _lib.ptr(dbd), _lib.ptr(dwc), _lib.ptr(dbc),
                                          1 if ctx.need_wgrad else 0, _lib.stream_ptr(out)), "head_backward")
        _lib.LAUNCHES += 1
        return None, dhs, None, None


def head_forward(model, hs):
    """The tail of RNNModel.forward (model_discCNN.py:547-566) on the [T, B, nhid] RNN output."""
    mask = None
    p = model.dropout_lin.p
    if model.training and p > 0.0:
        T, B, H = hs.shape
        # torch's own dropout draws the mask (same generator / Philox stream as the reference's nn.Dropout)
        mask = F.dropout(torch.ones(B, T * H, dtype=torch.float32, device=hs.device), p, True)
    return _HeadFunction.apply(model, hs, mask, model.decoder.weight)


def _loss_ws(dev):
    n = C.c_int64(0)
    _lib.check(_lib.lib().gnas_loss_workspace(C.byref(n)), "loss_workspace")
    return torch.empty(n.value, dtype=torch.float32, device=dev)


class _BCEFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, p, y):
        p = _lib.require(p.contiguous(), "probabilities")
        y = _lib.require(y.contiguous(), "targets")
        if p.shape != y.shape:
            raise ValueError(f"bce_loss: input {tuple(p.shape)} and target {tuple(y.shape)} differ")
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        ws = _loss_ws(p.device)
        _lib.check(_lib.lib().gnas_bce_forward(_lib.ptr(p), _lib.ptr(y), p.numel(), _lib.ptr(loss), _lib.ptr(ws),
                                               _lib.stream_ptr(p)), "bce_forward")
        _lib.LAUNCHES += 1
        ctx.save_for_backward(p, y)
        return loss

    @staticmethod
    def backward(ctx, gout):
        p, y = ctx.saved_tensors
        gout = _lib.require(gout.contiguous(), "grad_output")
        dp = torch.empty_like(p)
        _lib.check(_lib.lib().gnas_bce_backward(_lib.ptr(p), _lib.ptr(y), _lib.ptr(gout), p.numel(), _lib.ptr(dp),
                                                _lib.stream_ptr(p)), "bce_backward")
        _lib.LAUNCHES += 1
        return dp, None


class _TARFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, h):
        h = _lib.require(h.contiguous(), "rnn output")
        T = h.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=h.device)
        ws = _loss_ws(h.device)
        _lib.check(_lib.lib().gnas_tar_forward(_lib.ptr(h), T, h[0].numel(), _lib.ptr(loss), _lib.ptr(ws),
                                               _lib.stream_ptr(h)), "tar_forward")
        _lib.LAUNCHES += 1
        ctx.save_for_backward(h)
        return loss

    @staticmethod
    def backward(ctx, gout):
        h, = ctx.saved_tensors
        gout = _lib.require(gout.contiguous(), "grad_output")
        dh = torch.empty_like(h)
        _lib.check(_lib.lib().gnas_tar_backward(_lib.ptr(h), _lib.ptr(gout), h.shape[0], h[0].numel(), _lib.ptr(dh),
                                                _lib.stream_ptr(h)), "tar_backward")
        _lib.LAUNCHES += 1
        return dh


def bce_loss(p, y):
    """nn.BCELoss() (mean) of probabilities p against targets y (train_and_validate.py:136)."""
    return _BCEFunction.apply(p, y.to(p.dtype))


def tar_loss(rnn_h, beta):
    """Temporal activation regularisation beta * mean((h[1:] - h[:-1])^2), train_and_validate.py:140."""
    if rnn_h.shape[0] < 2:
        return beta * (rnn_h[1:] - rnn_h[:-1]).pow(2).mean()      # nan, as in the reference
    return beta * _TARFunction.apply(rnn_h)


def criterion_loss(criterion, p, y):
    """`criterion(p, y)`; a plain mean nn.BCELoss (what every reference driver passes) runs on the fused kernel."""
    if type(criterion) is torch.nn.BCELoss and criterion.reduction == "mean" and criterion.weight is None \
            and p.device.type == _lib.device_type() and p.dtype == torch.float32:
        return bce_loss(p, y)
    return criterion(p, y)
