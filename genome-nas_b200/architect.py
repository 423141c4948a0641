"""Architect: first-order alpha-gradient step of DARTS on the fused supernet.

Reference interface mirrored: darts_tools/architect.py:36-72 -- `Architect(model, criterion, args)` with
`args.arch_lr / arch_wdecay / wdecay / clip`, attribute `optimizer` (Adam over model.arch_parameters()),
`step(hidden_train, input_train, target_train, hidden_valid, input_valid, target_valid,
network_optimizer, unrolled) -> (hidden_next, None)`.

The reference's first-order step back-propagates the validation loss through the whole network and
then throws every weight gradient away (train_and_validate.py:112 zeroes them before they are
used).  Here the weight-gradient kernels are simply not launched for that pass; alpha gradients,
the Adam update and the returned hidden state are identical.
"""
import torch


class Architect(object):

    def __init__(self, model, criterion, args):
        self.network_weight_decay = args.wdecay
        self.network_clip = args.clip
        self.model = model
        self.optimizer = torch.optim.Adam(self.model.arch_parameters(), lr=args.arch_lr,
                                          weight_decay=args.arch_wdecay)
        self.criterion = criterion
        self.skip_weight_gradients = True

    def step(self, hidden_train, input_train, target_train, hidden_valid, input_valid, target_valid,
             network_optimizer, unrolled):
        if unrolled:
            raise NotImplementedError(
                "second-order (unrolled) architecture step is outside the accelerated path; "
                "genomeDARTS runs with unrolled=False (train_genomicDARTS.py default)")
        self.optimizer.zero_grad()
        hidden = self._backward_step(hidden_valid, input_valid, target_valid)
        self.optimizer.step()
        return hidden, None

    def _backward_step(self, hidden, input, target):
        toggle = self.skip_weight_gradients and hasattr(self.model, "set_weight_grads")
        if toggle:
            self.model.set_weight_grads(False)
        try:
            loss, hidden_next = self.model._loss(hidden, input, target, self.criterion)
            loss.backward()
        finally:
            if toggle:
                self.model.set_weight_grads(True)
        return hidden_next
