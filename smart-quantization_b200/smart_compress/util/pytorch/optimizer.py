"""OptimLP — the low-precision optimizer wrapper of reference
smart_compress/util/pytorch/optimizer.py:7-149, which is where weights, gradients and
SGD/Adam state meet the codec.

Order of one ``step(closure)`` (optimizer.py:129-143):
  closure() -> gradients compressed (tag optimizer_grad)           [_pre_closure,  :69-85]
  inner optimizer update
  gradients compressed AGAIN, weights (groups without ``no_weight_compression``), then
  ``momentum_buffer`` (SGD) or ``exp_avg`` / ``exp_avg_sq`` with all_positive=True (Adam/AdamW)
                                                                    [_post_closure, :87-127]
Results are assigned through ``.data`` so parameter identity is preserved.

When a quantizer exposes ``compress_many`` (the SmaQ plugin does: one multi-tensor launch for
all small tensors) whole phases are batched; otherwise tensors go one by one, as in the reference.
"""
from torch.optim import SGD, Adam, Optimizer
from torch.optim.adamw import AdamW

__all__ = ["OptimLP"]


def _skips(group, what):
    return bool(group.get(f"no_{what}_compression", False))


class OptimLP(Optimizer):
    def __init__(self, optim, weight_quant=None, grad_scaling=1.0, grad_quant=None, momentum_quant=None,
                 acc_quant=None):
        super().__init__(optim.param_groups, optim.defaults)  # placeholder init, state is shared below
        self.param_groups = optim.param_groups
        self.optim = optim

        assert grad_scaling > 0, "gradient scaling must be positive"
        self.grad_scaling = grad_scaling
        self.weight_quant = weight_quant
        self.grad_quant = grad_quant
        self.momentum_quant = momentum_quant
        self.acc_quant = acc_quant

        if isinstance(optim, SGD):
            self.momentum_keys = [("momentum_buffer", dict())]
        elif isinstance(optim, (Adam, AdamW)):
            self.momentum_keys = [("exp_avg", dict()), ("exp_avg_sq", dict(all_positive=True))]
        else:
            raise NotImplementedError("Only supporting Adam and SGD for now. ")

        if self.acc_quant is not None:
            self.weight_acc = {p: p.detach().clone().type_as(p) for g in self.param_groups for p in g["params"]}

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _apply(quant, holders, attr_get, attr_set, kwargs_list=None):
        """Run ``quant`` over many tensors; batched when the quantizer supports it."""
        if not holders:
            return
        tensors = [attr_get(h) for h in holders]
        many = getattr(quant, "compress_many", None)
        if many is not None:
            results = many(tensors, kwargs_list)
        elif kwargs_list is None:
            results = [quant(t) for t in tensors]
        else:
            results = [quant(t, **kw) for t, kw in zip(tensors, kwargs_list)]
        for h, r in zip(holders, results):
            attr_set(h, r)

    def _quantize_grads(self):
        if self.grad_quant is None:
            return
        params = [p for g in self.param_groups if not _skips(g, "grad")
                  for p in g["params"] if p.requires_grad and p.grad is not None]
        scale = self.grad_scaling

        def get(p):
            return p.grad.data if scale == 1.0 else p.grad.data * scale

        def put(p, r):
            p.grad.data = r.data

        self._apply(self.grad_quant, params, get, put)

    def _pre_closure(self):
        self._quantize_grads()
        if self.acc_quant is not None:  # switch accumulators in before stepping
            for g in self.param_groups:
                for p in g["params"]:
                    p.data = self.weight_acc[p].data

    def _post_closure(self):
        self._quantize_grads()

        if self.weight_quant is not None:
            params = [p for g in self.param_groups if not _skips(g, "weight") for p in g["params"]]
            self._apply(self.weight_quant, params, lambda p: p.data, lambda p, r: setattr(p, "data", r.data))

        if self.momentum_quant is not None:
            holders, kwargs_list = [], []
            for g in self.param_groups:
                if _skips(g, "momentum"):
                    continue
                if isinstance(self.optim, SGD) and g["momentum"] == 0:
                    continue
                for p in g["params"]:
                    if not p.requires_grad or p.grad is None:
                        continue
                    state = self.optim.state[p]
                    for key, kw in self.momentum_keys:
                        holders.append((state, key))
                        kwargs_list.append(kw)

            def put(h, r):
                h[0][h[1]].data = r.data

            self._apply(self.momentum_quant, holders, lambda h: h[0][h[1]], put, kwargs_list)

    def step(self, closure=None):
        """One update of the wrapped optimizer with compression before and after (a closure is required,
        as in the reference: the gradients are compressed right after it runs)."""

        def closure_(*args, **kwargs):
            value = closure(*args, **kwargs)
            self._pre_closure()
            return value

        loss = self.optim.step(closure=closure_)
        self._post_closure()
        return loss

    def __repr__(self):
        return "LP Optimizer: {}".format(self.optim.__repr__())

    def __str__(self):
        return "LP Optimizer: {}".format(self.optim.__str__())
