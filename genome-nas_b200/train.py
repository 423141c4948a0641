"""The supernet search step: alpha step + weight step.

`train()` keeps the reference's entry point (generalNAS_tools/train_and_validate.py:38-179) for the
genomeDARTS configuration (train_arch, first-order Architect or the P-DARTS inline alpha step).
`SearchStep` is the same arithmetic with the optimiser work done by the flat-buffer kernels
(gnas_grad_sumsq / gnas_sgd_step / gnas_adam_step) and an optional NCCL allreduce of the flat
gradient buffer for batch-sharded data parallelism; it is what bench.py times.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from .head import bce_loss, criterion_loss, tar_loss


class Hyper:
    """README.md:72 / train_genomicDARTS.py:53-140."""
    cnn_lr = 0.025
    cnn_wd = 3e-4
    momentum = 0.9
    rhn_lr = 8.0
    rhn_wd = 5e-7
    clip = 0.25
    beta = 1e-3
    arch_lr = 3e-3
    arch_wd = 1e-3
    adam_betas = (0.9, 0.999)
    adam_eps = 1e-8
    # P-DARTS / CWP-DARTS inline alpha step (train_and_validate.py:87-101, train_genomicPDARTS.py:300-301):
    # Adam(lr 6e-4, betas (0.5, 0.999), wd 1e-3) after clip_grad_norm_(arch_parameters, clip).  None = Architect
    # first-order step (no clipping of the alpha gradients).
    arch_clip = None




class SearchStep:
    """One genomeDARTS search iteration on the flat buffers of an RNNModelSearch."""

    def __init__(self, model, hyper=None, process_group=None, world_size=1):
        self.model, self.h = model, hyper or Hyper()
        self.pg, self.world = process_group, world_size
        self.region = model.flat_region()
        self.region.ensure_grads()
        off = self.region.offsets
        self.n_alpha = off["stem.0.weight"]                      # alphas_normal | alphas_reduce | weights (padded)
        self.rhn_lo, self.rhn_hi = off["rnns.0._W0"], off["decoder.weight"]
        self.total = self.region.total
        dev = self.region.flat.device
        self.mom = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(self.n_alpha, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(self.n_alpha, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.sumsq_a = torch.zeros(1, dtype=torch.float64, device=dev)
        self.t_adam = torch.zeros(1, dtype=torch.int32, device=dev)     # device-side so CUDA graphs stay valid
        self.graph = None
        self.mask_source = None     # optional callable(i, B, H, keep) -> [B, H] CPU mask (tests replay recorded masks)

    def _refresh(self):
        r = self.model.flat_region()
        if r.flat.data_ptr() != self.region.flat.data_ptr():
            raise RuntimeError("model parameters were re-allocated: build a new SearchStep")
        return r.flat, r.ensure_grads()

    def alpha_backward(self, x_valid, y_valid):
        """fwd + bwd of the validation BCE with the weight-gradient kernels switched off (architect.py:69-72)."""
        flat, g = self._refresh()
        model = self.model
        g[:self.n_alpha].zero_()
        model.set_weight_grads(False)
        try:
            logits, _ = model(x_valid, model.init_hidden(x_valid.size(0)))
            loss = bce_loss(logits, y_valid)
            loss.backward()
        finally:
            model.set_weight_grads(True)
        return loss.detach()

    def reduce_alpha_gradients(self):
        flat, g = self._refresh()
        if self.world > 1:
            ga = g[:self.n_alpha]
            torch.distributed.all_reduce(ga, group=self.pg)
            ga.div_(self.world)

    def alpha_update(self):
        """Adam on the three alpha tensors (architect.py:42,66)."""
        lib, h = _lib.lib(), self.h
        flat, g = self._refresh()
        st = _lib.stream_ptr(flat)
        sumsq, clip = None, 0.0
        if getattr(h, "arch_clip", None) is not None:
            self.sumsq_a.zero_()
            _lib.check(lib.gnas_grad_sumsq(_lib.ptr(g), self.n_alpha, _lib.ptr(self.sumsq_a), st), "grad_sumsq")
            sumsq, clip = self.sumsq_a, float(h.arch_clip)
        _lib.check(lib.gnas_adam_step(_lib.ptr(flat), _lib.ptr(g), _lib.ptr(self.adam_m), _lib.ptr(self.adam_v),
                                      self.n_alpha, h.arch_lr, h.arch_wd, h.adam_betas[0], h.adam_betas[1], h.adam_eps,
                                      0, _lib.ptr(self.t_adam), _lib.ptr(sumsq), clip, st), "adam_step")
        _lib.LAUNCHES += 1

    def alpha_step(self, x_valid, y_valid):
        loss = self.alpha_backward(x_valid, y_valid)
        self.reduce_alpha_gradients()
        self.alpha_update()
        return loss

    def weight_backward(self, x_train, y_train, zero=True):
        """fwd + bwd of BCE + beta*TAR on a training batch; gradients land in the flat buffer."""
        flat, g = self._refresh()
        model = self.model
        if zero:
            g.zero_()
        logits, hidden, rnn_hs, _ = model(x_train, model.init_hidden(x_train.size(0)), return_h=True)
        raw = bce_loss(logits, y_train)
        loss = raw + tar_loss(rnn_hs[-1], self.h.beta)
        loss.backward()
        return loss.detach(), raw.detach(), logits.detach()

    def reduce_gradients(self, scale=None):
        """batch-sharded data parallelism: mean of the per-rank gradients (one NCCL allreduce of 152 MB)."""
        flat, g = self._refresh()
        if self.world > 1:
            torch.distributed.all_reduce(g, group=self.pg)
            g.div_(self.world)
        elif scale is not None:
            g.mul_(scale)

    def apply_weight_update(self):
        """global-norm clip (train_and_validate.py:146-147) + SGD over both parameter groups
        (train_genomicDARTS.py:258-276: 'rnns' -> lr rhn_lr, no momentum; everything else, alphas included,
        lr cnn_lr with momentum)."""
        lib, h = _lib.lib(), self.h
        flat, g = self._refresh()
        st = _lib.stream_ptr(flat)
        self.sumsq.zero_()
        _lib.check(lib.gnas_grad_sumsq(_lib.ptr(g), self.total, _lib.ptr(self.sumsq), st), "grad_sumsq")
        first = 0     # momentum buffers start at zero: momentum*0 + d == d, torch's first-step rule
        for lo, hi in ((0, self.rhn_lo), (self.rhn_hi, self.total)):
            _lib.check(lib.gnas_sgd_step(_lib.ptr(flat[lo:]), _lib.ptr(g[lo:]), _lib.ptr(self.mom[lo:]), hi - lo,
                                         h.cnn_lr, h.cnn_wd, h.momentum, first, _lib.ptr(self.sumsq), h.clip, st),
                       "sgd_step")
        _lib.check(lib.gnas_sgd_step(_lib.ptr(flat[self.rhn_lo:]), _lib.ptr(g[self.rhn_lo:]), None,
                                     self.rhn_hi - self.rhn_lo, h.rhn_lr, h.rhn_wd, 0.0, first, _lib.ptr(self.sumsq),
                                     h.clip, st), "sgd_step")
        _lib.LAUNCHES += 4

    def weight_step(self, x_train, y_train):
        out = self.weight_backward(x_train, y_train)
        self.reduce_gradients()
        self.apply_weight_update()
        return out

    def _step_eager(self, x_train, y_train, x_valid, y_valid):
        la = self.alpha_step(x_valid, y_valid)
        lw, raw, logits = self.weight_step(x_train, y_train)
        return la, lw, raw, logits

    def step(self, x_train, y_train, x_valid, y_valid):
        if self.graph is None:
            return self._step_eager(x_train, y_train, x_valid, y_valid)
        for dst, src in zip(self._static_in, (x_train, y_train, x_valid, y_valid)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._draw_static_masks()
        if self.world == 1:
            self.graph.replay()
        else:                       # NCCL allreduces stay outside the graphs
            g1, g2, g3 = self.graph
            g1.replay()
            self.reduce_alpha_gradients()
            g2.replay()
            self.reduce_gradients()
            g3.replay()
        return self._static_out

    # ---- CUDA-graph mode: the ~3.7 k launches of a step are replayed from one captured graph --------------
    def _draw_static_masks(self):
        """Same CPU random stream as the eager path: x mask then h mask for the validation pass, then again
        for the training pass (utils.py:354-363 draws on the CPU generator)."""
        rnn = self.model.rnns[0]
        B, H = self._static_masks[0].shape
        for i, keep in enumerate((1. - rnn.dropoutx, 1. - rnn.dropouth) * 2):
            m = self.mask_source(i, B, H, keep) if self.mask_source is not None else torch.floor(torch.rand(B, H) + keep) / keep
            self._mask_host[i].copy_(m)
            self._static_masks[i].copy_(self._mask_host[i], non_blocking=True)

    def capture(self, x_train, y_train, x_valid, y_valid, warmup=2):
        """Capture alpha step + weight step into one CUDA graph.  `warmup` real steps run first on a side
        stream (they are ordinary optimisation steps on the given batch)."""
        dev = x_train.device
        rnn = self.model.rnns[0]
        B, H = x_train.size(0), rnn.nhid
        self._static_in = [t.clone() for t in (x_train, y_train, x_valid, y_valid)]
        self._static_masks = [torch.ones(B, H, device=dev) for _ in range(4)]
        self._mask_host = [torch.ones(B, H).pin_memory() for _ in range(4)]
        cursor = [0]

        def provider(b, h):
            i = cursor[0]
            cursor[0] = (i + 2) % 4
            return self._static_masks[i], self._static_masks[i + 1]

        rnn.__dict__["_mask_provider"] = provider
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._draw_static_masks()
                    self._step_eager(*self._static_in)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            cursor[0] = 0
            n0 = _lib.kernel_launches()
            xt, yt, xv, yv = self._static_in
            if self.world == 1:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._static_out = self._step_eager(xt, yt, xv, yv)
            else:
                pool = torch.cuda.graph_pool_handle()
                g1, g2, g3 = (torch.cuda.CUDAGraph() for _ in range(3))
                with torch.cuda.graph(g1, pool=pool):
                    la = self.alpha_backward(xv, yv)
                with torch.cuda.graph(g2, pool=pool):
                    self.alpha_update()
                    lw, raw, logits = self.weight_backward(xt, yt)
                with torch.cuda.graph(g3, pool=pool):
                    self.apply_weight_update()
                self._static_out = (la, lw, raw, logits)
                graph = (g1, g2, g3)
            self.graph_launches = _lib.kernel_launches() - n0      # our kernels inside one replay
            self.graph = graph
        except Exception:
            rnn.__dict__.pop("_mask_provider", None)
            self.graph = None
            raise
        return self


def train(train_queue, valid_queue, model, rhn, conv, criterion, optimizer, optimizer_a, architect, unrolled, lr,
          epoch, num_steps, clip_params, report_freq, beta, one_clip=True, train_arch=True, pdarts=True, task=None):
    """Drop-in for generalNAS_tools/train_and_validate.py:train with torch optimisers supplied by the driver.
    Returns (labels, predictions, mean loss) like the reference."""
    import numpy as np
    dev = next(model.parameters()).device
    labels, predictions, total, count = [], [], 0.0, 0
    valid_iter = None
    for step, (inp, target) in enumerate(train_queue):
        if step > num_steps:
            break
        model.train()
        inp, target = inp.float().to(dev), target.to(dev)
        bsz = inp.size(0)
        hidden = model.init_hidden(bsz)
        if train_arch:
            try:
                inp_s, tgt_s = next(valid_iter)
            except (StopIteration, TypeError):
                valid_iter = iter(valid_queue)
                inp_s, tgt_s = next(valid_iter)
            inp_s, tgt_s = inp_s.float().to(dev), tgt_s.to(dev)
            if pdarts:
                optimizer_a.zero_grad()
                logits, hidden, _, _ = model(inp_s, hidden, return_h=True)
                criterion_loss(criterion, logits, tgt_s).backward()
                torch.nn.utils.clip_grad_norm_(model.arch_parameters(), clip_params[2])
                optimizer_a.step()
            else:
                architect.step(hidden, inp, target, model.init_hidden(bsz), inp_s, tgt_s, optimizer, unrolled)
        optimizer.zero_grad()
        hidden = model.init_hidden(bsz)
        logits, hidden, rnn_hs, _ = model(inp, hidden, return_h=True)
        raw_loss = criterion_loss(criterion, logits, target)
        loss = raw_loss + sum(tar_loss(h, beta) for h in rnn_hs[-1:])
        loss.backward()
        if one_clip:
            torch.nn.utils.clip_grad_norm_(model.parameters(), clip_params[2])
        else:
            torch.nn.utils.clip_grad_norm_(conv, clip_params[0])
            torch.nn.utils.clip_grad_norm_(rhn, clip_params[1])
        optimizer.step()
        total += float(loss.detach()) * bsz
        count += bsz
        labels.append(target.detach().cpu().numpy())
        out = torch.softmax(logits, -1) if task == "next_character_prediction" else logits
        predictions.append(out.detach().cpu().numpy())
    return labels, predictions, np.float32(total / max(count, 1))


def infer(valid_queue, model, criterion, batch_size, num_steps, report_freq, task=None):
    """Drop-in for generalNAS_tools/train_and_validate.py:182-232: eval-mode (running-statistics BatchNorm, no
    dropout masks) forward over the validation queue; returns (labels, predictions, sample-weighted mean loss)."""
    import numpy as np
    dev = next(model.parameters()).device
    model.eval()
    labels, predictions, total, count = [], [], 0.0, 0
    for step, (inp, target) in enumerate(valid_queue):
        if step > num_steps:
            break
        inp, target = inp.to(dev).float(), target.to(dev)
        bsz = inp.size(0)
        with torch.no_grad():
            logits, _ = model(inp, model.init_hidden(bsz))
            loss = criterion_loss(criterion, logits, target)
        total += float(loss) * bsz
        count += bsz
        labels.append(target.detach().cpu().numpy())
        out = torch.softmax(logits, -1) if task == "next_character_prediction" else logits
        predictions.append(out.detach().cpu().numpy())
    return labels, predictions, np.float32(total / max(count, 1))
