"""Fused clip_grad_norm_ + Adam over a FlatModule's flat buffers (K12), with a torch.optim.Adam
compatible surface (param_groups[0]['lr'], state_dict()/load_state_dict(), step(), zero_grad()).
Replaces reference stage2_accompaniment/train.py:79-81,318-326 / stage1_compose/train.py:63-65,287-304."""
import math
import torch

from . import ops


class FusedAdam:
    def __init__(self, model, lr, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.0, grad_sync=None):
        self.model = model
        self.grad_sync = grad_sync        # dp.GradSync: one all-reduce of the flat gradient before the clip
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False,
                              "initial_lr": lr, "params": list(range(len(list(model.parameters()))))}]
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self.grad_scale = 1.0
        self._step = 0
        self._m = torch.zeros_like(model._flat)
        self._v = torch.zeros_like(model._flat)
        self._gn = torch.zeros(1, dtype=torch.float32, device=model._flat.device)
        self.last_grad_norm_sq = self._gn

    def zero_grad(self, set_to_none=False):
        self.model.zero_grad()

    def step(self):
        m = self.model
        if self._m.device != m._flat.device:
            self._m, self._v, self._gn = self._m.to(m._flat.device), self._v.to(m._flat.device), self._gn.to(m._flat.device)
        if self.grad_sync is not None:
            self.grad_sync.allreduce_grads()
        g = self.param_groups[0]
        self._step += 1
        gn = None
        if self.max_grad_norm > 0:
            self._gn.zero_()
            ops.sumsq(m._flat_grad, self._gn)
            gn = self._gn
        lp = None
        if m.compute_dtype == torch.bfloat16:
            m.weights()               # make sure the shadow buffer exists
            lp = m._flat_lp
        ops.adam_step(m._flat, m._flat_grad, self._m, self._v, lp, float(g["lr"]), g["betas"][0], g["betas"][1],
                      g["eps"], self._step, gn, self.max_grad_norm, self.grad_scale, zero_grad=True)
        if lp is not None:
            m.mark_lp_fresh()

    # ---- torch.optim.Adam-format checkpoints (reference optim/ep*_optim.pt) ------------------------
    def state_dict(self):
        """indexed like torch.optim.Adam over the REFERENCE module's parameters() (FlatModule.reference_param_names):
        a reference optim/ep*_optim.pt resumes here and a checkpoint written here resumes there"""
        state = {}
        for name, shape, off, n, idx in self._ref_layout():
            state[idx] = {"step": torch.tensor(float(self._step)),
                          "exp_avg": self._m[off:off + n].view(shape).clone(),
                          "exp_avg_sq": self._v[off:off + n].view(shape).clone()}
        pg = dict(self.param_groups[0])
        return {"state": dict(sorted(state.items())), "param_groups": [pg]}

    def _ref_layout(self):
        order = {name: i for i, name in enumerate(self.model.reference_param_names())}
        return [(name, shape, off, n, order[name]) for name, shape, off, n, _ in self.model._specs]

    def load_state_dict(self, sd):
        for name, shape, off, n, idx in self._ref_layout():
            st = sd["state"].get(idx)
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(shape):
                raise ValueError("optimizer state %d has shape %s, parameter %s has %s: not a checkpoint of this "
                                 "architecture" % (idx, tuple(st["exp_avg"].shape), name, tuple(shape)))
            self._m[off:off + n].view(shape).copy_(st["exp_avg"])
            self._v[off:off + n].view(shape).copy_(st["exp_avg_sq"])
            self._step = int(float(st["step"]))
        for k in ("lr", "betas", "eps"):
            if k in sd["param_groups"][0]:
                self.param_groups[0][k] = sd["param_groups"][0][k]


class WarmupCosine:
    """LR schedule of the reference loops: linear warm-up written into param_groups[0]['lr'], then
    CosineAnnealingLR stepped with an explicit epoch (closed form), train.py:100-104."""

    def __init__(self, optimizer, max_lr, min_lr, warmup_steps, t_max, accum_steps=1):
        self.opt, self.max_lr, self.min_lr = optimizer, max_lr, min_lr
        self.warmup, self.t_max, self.accum = warmup_steps, t_max, accum_steps

    def update(self, train_steps):
        if (train_steps // self.accum) < self.warmup:
            lr = self.max_lr * train_steps / (self.warmup * self.accum)
        else:
            t = train_steps // self.accum - self.warmup
            lr = self.min_lr + (self.max_lr - self.min_lr) * (1 + math.cos(math.pi * t / self.t_max)) / 2
        self.opt.param_groups[0]["lr"] = lr
        return lr
