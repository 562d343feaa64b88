"""Shared front/back end of the two stage-2 models: embedding (+segment, +PE), output projection,
cross-entropy, the fused train step and the parameter layout."""
import math
import torch

from .. import ops
from ..engine import FlatModule, ModelFn, CrossEntropyFn, site_seed

D_MODEL = 512
N_HEAD = 8
MAX_POS = 12000


def _normal(std, mean=0.0):
    def init(t):
        t.normal_(mean, std)
    return init


def _zeros(t):
    t.zero_()


def sinusoid_pe(max_pos, d):
    # reference transformer_helpers.py:48-54
    pe = torch.zeros(max_pos, d)
    position = torch.arange(0, max_pos, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(1)


class Stage2Base(FlatModule):
    def __init__(self, n_token, n_layer, n_head, d_model, d_ff, d_embed, activation, dropout, use_pe,
                 use_segment_emb, n_segment_types, use_chord_mhot_emb, compute_dtype):
        super().__init__(compute_dtype)
        if d_model != D_MODEL or n_head != N_HEAD or d_embed != d_model:
            raise ValueError("the B200 kernels are built for d_model = d_embed = 512 and 8 heads "
                             "(the reference configs); got d_model=%d n_head=%d d_embed=%d" % (d_model, n_head, d_embed))
        if use_chord_mhot_emb:
            raise NotImplementedError("use_chord_mhot_emb is never enabled by the reference scripts (train.py:293)")
        if d_ff % 64 != 0:
            raise ValueError("d_ff must be a multiple of 64")
        self.n_token, self.n_layer, self.n_head = n_token, n_layer, n_head
        self.d_model, self.d_ff, self.d_embed = d_model, d_ff, d_embed
        self.dropout, self.activation, self.use_pe = dropout, activation, use_pe
        self.use_segment_emb = bool(use_segment_emb)
        self.n_segment_types = n_segment_types
        self.use_chord_mhot_emb = False
        self.ldv = (n_token + 7) // 8 * 8          # padded logits / dlogits leading dim
        self._add_param("token_emb.emb_lookup.weight", (n_token, d_embed), _normal(0.01))
        if self.use_segment_emb:
            self._add_param("segemb.emb_lookup.weight", (n_segment_types, d_embed), _normal(0.01))
        self._add_param("dec_out_proj.weight", (n_token, d_model), _normal(0.01))
        self._add_param("dec_out_proj.bias", (n_token,), _zeros)

    def _finish(self):
        self._finalize()
        self._add_buffer("pe.pe", sinusoid_pe(MAX_POS, self.d_embed))
        if not self.use_segment_emb:
            self.segemb = None
        self._sl = self._slices()
        print('[info] model init completed')

    # ---- helpers ---------------------------------------------------------------------------
    def _ref_order_key(self, name, index):
        # reference modules create `segemb` last (music_performer.py:23-47, music_gpt2.py:33-62)
        return (1 if name.startswith("segemb.") else 0, index)

    def _wv(self, buf, name):
        # views of the flat buffers are cached per (buffer address, name): slicing + view is two torch dispatches
        # (~5 us), ~430 of them per train step -- a third of the host time of a step at the reference's batch size 4
        cache = self.__dict__.setdefault("_view_cache", {})
        key = (buf.data_ptr(), name)
        v = cache.get(key)
        if v is None:
            off, n, shape = self._sl[name]
            v = buf[off:off + n].view(shape)
            if len(cache) > 8192:          # buffers were re-allocated many times (.to / .cuda): drop the stale views
                cache.clear()
            cache[key] = v
        return v

    def _gv(self, name):
        return self._wv(self._flat_grad, name)

    def _p_drop(self):
        return float(self.dropout) if self.training else 0.0

    # ---- reference-facing API ----------------------------------------------------------------
    def forward(self, x, seg_inp=None, chord_inp=None, keep_last_only=False, attn_kwargs=None):
        if not x.is_cuda:
            raise RuntimeError("emo_disentanger_b200 models run on CUDA only (no CPU fallback); call .cuda()")
        if not (seg_inp is not None and self.use_segment_emb):
            seg_inp = None
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            anchor = self._parameters_anchor()
            logits = ModelFn.apply(anchor, self, (x, seg_inp))
        else:
            logits, _ = self._forward_impl(x, seg_inp, save=False)
        if keep_last_only:
            logits = logits[:, -1, :]
        return logits

    def _parameters_anchor(self):
        return self._wv(self._flat, "dec_out_proj.bias") if False else dict(self.named_parameters())["dec_out_proj.bias"]

    def compute_loss(self, dec_logits, dec_tgt, reduction='mean'):
        if reduction != 'mean':
            raise NotImplementedError("only reduction='mean' is used by the reference")
        recons_loss = CrossEntropyFn.apply(dec_logits, dec_tgt, self.n_token - 1, True)
        return {'recons_loss': recons_loss, 'total_loss': recons_loss}

    # ---- fused train step (no autograd): fwd + CE + bwd, grads accumulated in _flat_grad -------
    def train_step(self, x, seg_inp, dec_tgt, gscale=1.0, count_allreduce=None):
        """Returns a 3-element fp32 tensor [n_valid, loss_sum, n_correct]; gradients of
        gscale * mean-CE are accumulated into the flat gradient buffer."""
        if not (seg_inp is not None and self.use_segment_emb):
            seg_inp = None
        self._prepare_grads()
        hid, saved = self._forward_hidden(x, seg_inp, save=True)
        R = hid.shape[0]
        logits = self._logits(hid)
        acc = torch.zeros(3, dtype=torch.float32, device=x.device)
        ops.ce_count(dec_tgt, self.n_token - 1, acc[0:1])
        count = acc[0:1]
        if count_allreduce is not None:
            count = count_allreduce(acc[0:1])
        dl = torch.empty(R, self.ldv, dtype=self.compute_dtype, device=x.device)
        ops.ce_fwd_bwd(logits, dec_tgt, self.n_token, self.n_token - 1, count, acc[1:2], acc[2:3], None, dl, gscale)
        self._backward_from_dl(saved, hid, dl)
        return acc

    # ---- forward / backward plumbing shared by the subclasses ---------------------------------
    def _embed(self, x, seg, seed):
        B, T = x.shape
        W = self._flat
        h = torch.empty(B * T, self.d_model, dtype=self.compute_dtype, device=x.device)
        ops.embed_fwd(x, seg, self._wv(W, "token_emb.emb_lookup.weight"),
                      self._wv(W, "segemb.emb_lookup.weight") if seg is not None else None,
                      self.pe.pe if self.use_pe else None, h, self.d_model ** 0.5, self._p_drop(), site_seed(seed, 0))
        return h

    def _logits(self, hid):
        Wc = self.weights()
        R = hid.shape[0]
        logits = torch.empty(R, self.ldv, dtype=torch.float32, device=hid.device)
        ops.linear_fwd(hid, self._wv(Wc, "dec_out_proj.weight"), logits[:, :self.n_token],
                       bias=self._wv(self._flat, "dec_out_proj.bias"))
        return logits

    def _forward_impl(self, x, seg, save):
        B, T = x.shape
        hid, saved = self._forward_hidden(x, seg, save)
        logits = self._logits(hid)
        out = logits[:, :self.n_token].view(B, T, self.n_token)
        if save:
            saved["hid"] = hid
        return out, saved

    def _backward_impl(self, saved, dlogits):
        """autograd path: dlogits [B,T,V] fp32 (any strides) -> padded compute-dtype buffer."""
        B, T, V = dlogits.shape
        dl = torch.zeros(B * T, self.ldv, dtype=self.compute_dtype, device=dlogits.device)
        dl[:, :V].copy_(dlogits.reshape(B * T, V))      # layout/dtype plumbing only
        self._backward_from_dl(saved, saved["hid"], dl)

    def _backward_from_dl(self, saved, hid, dl):
        Wc = self.weights()
        V = self.n_token
        R = hid.shape[0]
        dlv = dl[:, :V]
        ops.linear_wgrad(dlv, hid, self._gv("dec_out_proj.weight"))
        ops.colsum(dl, self._gv("dec_out_proj.bias"), n=V)
        dh = torch.empty(R, self.d_model, dtype=self.compute_dtype, device=dl.device)
        ops.linear_dgrad(dlv, self._wv(Wc, "dec_out_proj.weight"), dh)
        dx = self._backward_hidden(saved, dh)
        x, seg, seed, p = saved["tokens"], saved["seg"], saved["seed"], saved["p"]
        ops.embed_bwd(x, seg, dx, self._gv("token_emb.emb_lookup.weight"),
                      self._gv("segemb.emb_lookup.weight") if seg is not None else None,
                      self.d_model ** 0.5, p, site_seed(seed, 0))
