"""Host-side plumbing shared by the three model modules: flat fp32 parameter / gradient buffers
(one NCCL all-reduce, one fused clip+Adam launch), bf16 shadow weights, dropout seeds, and the
autograd glue that lets the reference's `loss.backward()` drive the hand-written backward."""
import torch
from torch import nn

from . import ops

_MASK64 = (1 << 64) - 1
ALIGN = 8  # elements: keeps every bf16 shadow weight 16-byte aligned (TMA requirement)


def site_seed(base, site):
    x = (base + (site + 1) * 0x9E3779B97F4A7C15) & _MASK64
    x ^= x >> 31
    x = (x * 0xBF58476D1CE4E5B9) & _MASK64
    x ^= x >> 29
    return x & _MASK64


class FlatModule(nn.Module):
    """nn.Module whose parameters are views into ONE flat fp32 buffer (`_flat`), with gradients in
    `_flat_grad` and (bf16 mode) shadow weights in `_flat_lp`.  state_dict keys are whatever nested
    names were registered, so reference checkpoints load unchanged."""

    def __init__(self, compute_dtype=torch.bfloat16):
        super().__init__()
        if compute_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("compute_dtype must be torch.bfloat16 or torch.float32")
        self.compute_dtype = compute_dtype
        self._specs = []          # (dotted name, shape, offset, numel)
        self._total = 0
        self._flat = None
        self._flat_grad = None
        self._flat_lp = None
        self._lp_version = -1
        self._seed_base = torch.initial_seed() & _MASK64
        self._fwd_calls = 0

    # ---- registration ----------------------------------------------------------------------
    def _container(self, path):
        mod = self
        for part in path:
            if part not in mod._modules:
                mod.add_module(part, nn.Module())
            mod = mod._modules[part]
        return mod

    def _add_param(self, name, shape, init):
        n = 1
        for s in shape:
            n *= s
        off = self._total
        self._specs.append((name, tuple(shape), off, n, init))
        self._total = off + (n + ALIGN - 1) // ALIGN * ALIGN

    def _add_buffer(self, name, tensor, persistent=True):
        parts = name.split(".")
        self._container(parts[:-1]).register_buffer(parts[-1], tensor, persistent=persistent)

    def _finalize(self, device="cpu"):
        flat = torch.zeros(self._total, dtype=torch.float32, device=device)
        for name, shape, off, n, init in self._specs:
            view = flat[off:off + n].view(shape)
            init(view)
            parts = name.split(".")
            self._container(parts[:-1]).register_parameter(parts[-1], nn.Parameter(view))
        self._bind(flat)

    # ---- reference parameter order ------------------------------------------------------------
    def _ref_order_key(self, name, index):
        """sort key that turns the registration order into the order of the REFERENCE module's parameters()
        (torch.optim state dicts are indexed by it); subclasses override where the two differ"""
        return index

    def reference_param_names(self):
        names = [name for name, _ in self.named_parameters()]     # module-tree order, as torch enumerates them
        return [n for _, n in sorted((self._ref_order_key(n, i), n) for i, n in enumerate(names))]

    def _named_flat_params(self):
        # cached: walking named_parameters() on every backward was ~1 ms of host time per step (the Parameter objects are
        # stable -- .to() / .cuda() rebind their .data in _bind; _apply drops the cache before it re-reads them)
        ps = self.__dict__.get("_nfp_cache")
        if ps is None:
            d = dict(self.named_parameters())
            ps = [(d[name], shape, off, n) for name, shape, off, n, _ in self._specs]
            self.__dict__["_nfp_cache"] = ps
        return ps

    def _bind(self, flat):
        self._flat = flat
        self._flat_grad = torch.zeros_like(flat)
        self._flat_lp = None
        self._lp_version = -1
        self.__dict__.pop("_version_params", None)
        self.__dict__.pop("_view_cache", None)            # cached views (_wv) alias the buffers that were just replaced
        for p, shape, off, n in self._named_flat_params():
            p.data = flat[off:off + n].view(shape)
            p.grad = self._flat_grad[off:off + n].view(shape)

    def _apply(self, fn, recurse=True):
        super()._apply(fn)
        self.__dict__.pop("_nfp_cache", None)
        ps = self._named_flat_params()
        if ps:
            dev = ps[0][0].device
            flat = torch.zeros(self._total, dtype=torch.float32, device=dev)
            for p, shape, off, n in ps:
                flat[off:off + n].copy_(p.data.reshape(-1).float())
            self._bind(flat)
        return self

    # ---- gradient buffer -------------------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        self._flat_grad.zero_()
        self._attach_grads()

    def _attach_grads(self):
        for p, shape, off, n in self._named_flat_params():
            g = p.grad
            if g is None or g.data_ptr() != self._flat_grad.data_ptr() + off * 4:
                p.grad = self._flat_grad[off:off + n].view(shape)

    def _prepare_grads(self):
        """Called at the start of backward: if an external optimizer dropped the .grad views
        (zero_grad(set_to_none=True)), start from a zeroed flat buffer and re-attach them."""
        ps = self._named_flat_params()
        if any(p.grad is None for p, _, _, _ in ps):
            self._flat_grad.zero_()
        self._attach_grads()

    def grad_range(self, prefix):
        """[lo, hi) of the flat gradient buffer that holds every parameter whose name starts with `prefix`
        (registration keeps a layer's parameters contiguous): the unit of the bucketed gradient all-reduce"""
        sel = [(off, off + n) for name, shape, off, n, _ in self._specs if name.startswith(prefix)]
        if not sel:
            raise KeyError(prefix)
        lo, hi = min(a for a, _ in sel), max(b for _, b in sel)
        inside = sum(1 for name, shape, off, n, _ in self._specs if lo <= off < hi)
        if inside != len(sel):
            raise RuntimeError("parameters of %r are not contiguous in the flat buffer" % prefix)
        return lo, hi

    def _layer_done(self, l):
        """called by the backward of each model after the last kernel that writes layer l's gradients was launched"""
        hook = self.__dict__.get("grad_hook")
        if hook is not None:
            hook(l)

    def grad_view(self, p_name):
        for name, shape, off, n, _ in self._specs:
            if name == p_name:
                return self._flat_grad[off:off + n].view(shape)
        raise KeyError(p_name)

    # ---- compute-dtype weights -----------------------------------------------------------
    def weights(self):
        """Flat buffer in the compute dtype (bf16 shadow refreshed when the fp32 masters changed)."""
        if self.compute_dtype == torch.float32:
            return self._flat
        ver = self._weights_version()
        if self._flat_lp is None or self._flat_lp.device != self._flat.device:
            self._flat_lp = torch.empty(self._total, dtype=torch.bfloat16, device=self._flat.device)
            self.__dict__.pop("_view_cache", None)
            self._lp_version = -1
        if ver != self._lp_version:
            ops.cast(self._flat, self._flat_lp)
            self._lp_version = ver
        return self._flat_lp

    def _weights_version(self):
        """Changes whenever the fp32 masters may have changed.  `_flat._version` alone is not enough: every
        Parameter is a view bound with `p.data = ...`, which gives it its OWN version counter, so an external
        optimizer (`torch.optim.Adam.step`) or `load_state_dict` writes `_flat`'s memory without moving
        `_flat._version`.  The counters only grow, so their sum is a valid key."""
        ps = self.__dict__.get("_version_params")
        if ps is None:
            ps = [p for p, _, _, _ in self._named_flat_params()]
            self.__dict__["_version_params"] = ps
        v = self._flat._version
        for p in ps:
            v += p._version
        return v

    def mark_lp_fresh(self):
        """The fused optimizer wrote the shadow weights itself."""
        self._lp_version = self._weights_version()

    def w(self, flatbuf, name_to_slice):
        off, n, shape = name_to_slice
        return flatbuf[off:off + n].view(shape)

    def _slices(self):
        return {name: (off, n, shape) for name, shape, off, n, _ in self._specs}

    def next_seed(self):
        self._fwd_calls += 1
        return site_seed(self._seed_base, self._fwd_calls)


class ModelFn(torch.autograd.Function):
    """Autograd glue: forward runs the CUDA forward and keeps the saved activations; backward runs
    the CUDA backward, which accumulates straight into the flat gradient buffer."""

    @staticmethod
    def forward(ctx, anchor, model, args):
        out, saved = model._forward_impl(*args, save=True)
        ctx.model = model
        ctx.saved = saved
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.model._prepare_grads()
        ctx.model._backward_impl(ctx.saved, dout)
        ctx.saved = None
        return None, None, None


class CrossEntropyFn(torch.autograd.Function):
    """mean CE over targets != ignore_index on device; backward hands (softmax - onehot)/count."""

    @staticmethod
    def forward(ctx, logits, tgt, ignore_index, batch_first):
        V = logits.shape[-1]
        l2 = logits.reshape(-1, V) if batch_first else logits.transpose(0, 1).reshape(-1, V)
        if l2.dtype != torch.float32 or l2.stride(-1) != 1:
            l2 = l2.float().contiguous()
        dev = logits.device
        acc = torch.zeros(3, dtype=torch.float32, device=dev)   # count, loss_sum, ncorrect
        ops.ce_count(tgt, ignore_index, acc[0:1], batch_first)
        dl = torch.empty(l2.shape[0], V, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        ops.ce_fwd_bwd(l2, tgt, V, ignore_index, acc[0:1], acc[1:2], acc[2:3], None, dl, 1.0, batch_first)
        ctx.dl = dl
        ctx.shape = logits.shape
        ctx.batch_first = batch_first
        return acc[1] / acc[0]

    @staticmethod
    def backward(ctx, g):
        dl = ctx.dl
        if ctx.batch_first:
            d = dl.view(ctx.shape)
        else:
            T, B, V = ctx.shape
            d = dl.view(B, T, V).transpose(0, 1)
        return d * g, None, None, None
