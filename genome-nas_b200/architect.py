"""Architect: first-order and second-order (unrolled) alpha-gradient steps of DARTS on the fused supernet.

Reference interface mirrored: darts_tools/architect.py:36-72 -- `Architect(model, criterion, args)` with
`args.arch_lr / arch_wdecay / wdecay / clip`, attribute `optimizer` (Adam over model.arch_parameters()),
`step(hidden_train, input_train, target_train, hidden_valid, input_valid, target_valid,
network_optimizer, unrolled) -> (hidden_next, None)`.

The reference's first-order step back-propagates the validation loss through the whole network and
then throws every weight gradient away (train_and_validate.py:112 zeroes them before they are
used).  Here the weight-gradient kernels are simply not launched for that pass; alpha gradients,
the Adam update and the returned hidden state are identical.

The unrolled step (architect.py:74-129) is the reference's arithmetic on the same kernels: one virtual SGD
step w' = w - eta (clip(dw L_train) + wd w) on a copy of the model, dalpha L_val(w', alpha), and the implicit
term by the finite-difference Hessian-vector product (two more passes at w +- R dw' L_val, r = 1e-2).  The fused
modules deliver weight gradients into `p.grad` (views of one flat buffer) instead of returning them through
autograd, so the weight gradients are taken with backward() where the reference calls torch.autograd.grad;
the alpha gradients of the two finite-difference passes need no weight-gradient kernel at all.
"""
import torch


def _concat(xs):
    return torch.cat([x.reshape(-1) for x in xs])


def _clip(grads, max_norm):
    """architect.py:22-32: in-place global-norm clip, returns the coefficient (a 0-d tensor)"""
    total_norm = sum(g.data.norm(2) ** 2 for g in grads) ** 0.5
    clip_coef = max_norm / (total_norm + 1e-6)
    if clip_coef < 1:
        for g in grads:
            g.data.mul_(clip_coef)
    return clip_coef


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
        eta = network_optimizer.param_groups[0]['lr']
        self.optimizer.zero_grad()
        if unrolled:
            hidden = self._backward_step_unrolled(hidden_train, input_train, target_train, hidden_valid, input_valid,
                                                  target_valid, eta)
        else:
            hidden = self._backward_step(hidden_valid, input_valid, target_valid)
        self.optimizer.step()
        return hidden, None

    # ---- second order (architect.py:45-53, 74-129) -------------------------------------------------------
    @staticmethod
    def _weight_grads(model, loss):
        """d loss / d every parameter (alphas included), as fresh tensors"""
        for p in model.parameters():
            p.grad = None
        loss.backward()
        return [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]

    def _compute_unrolled_model(self, hidden, input, target, eta):
        loss, hidden_next = self.model._loss(hidden, input, target, self.criterion)
        theta = _concat(self.model.parameters()).data.clone()
        grads = self._weight_grads(self.model, loss)
        clip_coef = _clip(grads, self.network_clip)
        dtheta = _concat(grads).data + self.network_weight_decay * theta
        return self._construct_model_from_theta(theta.sub_(dtheta, alpha=eta)), clip_coef

    def _construct_model_from_theta(self, theta):
        model_new = self.model.new()
        model_dict = self.model.state_dict()
        params, offset = {}, 0
        for k, v in self.model.named_parameters():
            n = v.numel()
            params[k] = theta[offset:offset + n].view(v.size())
            offset += n
        assert offset == len(theta)
        model_dict.update(params)
        model_new.load_state_dict(model_dict)
        return model_new.to(next(self.model.parameters()).device)

    def _backward_step_unrolled(self, hidden_train, input_train, target_train, hidden_valid, input_valid,
                                target_valid, eta):
        unrolled_model, clip_coef = self._compute_unrolled_model(hidden_train, input_train, target_train, eta)
        unrolled_model.train(self.model.training)
        unrolled_loss, hidden_next = unrolled_model._loss(hidden_valid, input_valid, target_valid, self.criterion)
        for p in unrolled_model.parameters():
            p.grad = None
        unrolled_loss.backward()
        for p in unrolled_model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        dalpha = [v.grad for v in unrolled_model.arch_parameters()]
        dtheta = [v.grad for v in unrolled_model.parameters()]      # the alphas are parameters too: clipped along
        _clip(dtheta, self.network_clip)
        vector = [dt.data.clone() for dt in dtheta]
        implicit_grads = self._hessian_vector_product(vector, hidden_train, input_train, target_train, r=1e-2)
        for g, ig in zip(dalpha, implicit_grads):
            g.data.sub_(ig.data * (eta * clip_coef))
        for v, g in zip(self.model.arch_parameters(), dalpha):
            if v.grad is None:
                v.grad = g.data.clone()
            else:
                v.grad.data.copy_(g.data)
        return hidden_next

    def _hessian_vector_product(self, vector, hidden, input, target, r=1e-2):
        R = r / _concat(vector).norm()
        toggle = hasattr(self.model, "set_weight_grads")
        if toggle:
            self.model.set_weight_grads(False)       # only d/dalpha is needed from the two perturbed passes
        try:
            for p, v in zip(self.model.parameters(), vector):
                p.data.add_(v, alpha=float(R))
            loss, _ = self.model._loss(hidden, input, target, self.criterion)
            grads_p = torch.autograd.grad(loss, self.model.arch_parameters())
            for p, v in zip(self.model.parameters(), vector):
                p.data.sub_(v, alpha=2 * float(R))
            loss, _ = self.model._loss(hidden, input, target, self.criterion)
            grads_n = torch.autograd.grad(loss, self.model.arch_parameters())
            for p, v in zip(self.model.parameters(), vector):
                p.data.add_(v, alpha=float(R))
        finally:
            if toggle:
                self.model.set_weight_grads(True)
        return [(x - y).div_(2 * R) for x, y in zip(grads_p, grads_n)]

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
