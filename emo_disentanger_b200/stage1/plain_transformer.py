"""PlainTransformer -- drop-in for reference stage1_compose/model/plain_transformer.py (+ the live subset
of optimus_txl_decoder.py: attn_type 0, pre-LN, no segment embedding, no cross attention), computing on
hand-written sm_100a kernels through libemo_b200.so.

Same constructor / forward / generate / compute_loss signatures, `[T, B]` token layout and the same 138
state-dict entries as the reference.  Internally activations are batch-major `[B, T, d]`
(the attention kernels walk one sequence at a time); tensors that cross the API (`logits [T,B,V]`, `mems
[m,B,d]`) are transposed views.

Per layer (RelPartialLearnableDecoderLayer, optimus_txl_decoder.py:305-387, 47-61):
  a     = LN(cat[mem, h])
  heads = a Wqkv^T                          tcgen05 GEMM (no bias)
  r     = drop(sinusoid(klen-1..0)) Wr^T    GEMM on the [klen, 512] position table
  att   = rel-pos attention                 attn.cu (REL): AC + shifted BD, softmax, dropatt + renormalise
  h1    = h + drop(att Wo^T)                GEMM epilogue: dropout + residual
  c     = LN(h1);  f = drop(relu(c W1^T + b1));  h2 = h1 + drop(f W2^T + b2)
After the last layer only dropout (no final LN, :918) and `dec_out_proj`.
Memory = the n_layer+1 layer INPUTS (hidden states), kept to the last `mem_len` rows (:702-722);
training uses mem_len = 0 (YAML), decoding mem_len = tgt_len = 512 (inference.py:180-181).
"""
import torch

from .. import ops
from ..engine import FlatModule, ModelFn, CrossEntropyFn, site_seed

E = 64


def _normal(std, mean=0.0):
    def init(t):
        t.normal_(mean, std)
    return init


def _zeros(t):
    t.zero_()


class PlainTransformer(FlatModule):
    def __init__(self, d_word_embed, vocab_size, dec_n_layer, dec_n_head, dec_d_model, dec_d_ff,
                 dec_mem_len, dec_tgt_len, dec_dropout=0.1, dec_activation='relu', pad_index=None,
                 pre_lnorm=False, compute_dtype=torch.bfloat16):
        super().__init__(compute_dtype)
        if dec_d_model != 512 or dec_n_head != 8 or d_word_embed != dec_d_model:
            raise ValueError("the B200 kernels are built for d_model = d_word_embed = 512 and 8 heads")
        if not pre_lnorm:
            raise NotImplementedError("every reference config sets pre_lnorm: True (stage1_compose/config/*.yaml)")
        if dec_activation != 'relu':
            raise NotImplementedError("PositionwiseFF is ReLU-only in the reference (optimus_txl_decoder.py:37)")
        self.d_word_embed, self.vocab_size = d_word_embed, vocab_size
        self.dec_n_layer, self.dec_n_head, self.dec_d_model, self.dec_d_ff = dec_n_layer, dec_n_head, dec_d_model, dec_d_ff
        self.dec_dropout, self.dec_activation = dec_dropout, dec_activation
        self.dec_mem_len, self.dec_tgt_len = dec_mem_len, dec_tgt_len
        self.pad_index = vocab_size - 1 if pad_index is None else pad_index
        self.ldv = (vocab_size + 7) // 8 * 8
        d, f, V = dec_d_model, dec_d_ff, vocab_size

        def emb_init(t):
            t.normal_(0.0, 0.01)          # weights_init re-draws the whole table, padding row included
        self._add_param("word_emb.emb_lookup.weight", (V, d), emb_init)
        self._add_param("decoder.r_w_bias", (dec_n_head, E), _normal(0.01))
        self._add_param("decoder.r_r_bias", (dec_n_head, E), _normal(0.01))
        for l in range(dec_n_layer):
            p = "decoder.layers.%d." % l
            self._add_param(p + "dec_attn.qkv_net.weight", (3 * d, d), _normal(0.01))
            self._add_param(p + "dec_attn.r_net.weight", (d, d), _normal(0.01))
            self._add_param(p + "dec_attn.o_net.weight", (d, d), _normal(0.01))
            self._add_param(p + "dec_attn.layer_norm.weight", (d,), _normal(0.01, 1.0))
            self._add_param(p + "dec_attn.layer_norm.bias", (d,), _zeros)
            self._add_param(p + "pos_ff.CoreNet.0.weight", (f, d), _normal(0.01))
            self._add_param(p + "pos_ff.CoreNet.0.bias", (f,), _zeros)
            self._add_param(p + "pos_ff.CoreNet.3.weight", (d, f), _normal(0.01))
            self._add_param(p + "pos_ff.CoreNet.3.bias", (d,), _zeros)
            self._add_param(p + "pos_ff.layer_norm.weight", (d,), _normal(0.01, 1.0))
            self._add_param(p + "pos_ff.layer_norm.bias", (d,), _zeros)
        self._add_param("dec_out_proj.weight", (V, d), _normal(0.01))
        self._add_param("dec_out_proj.bias", (V,), _zeros)
        self._finalize()
        inv_freq = 1 / (10000 ** (torch.arange(0.0, d, 2.0) / d))
        self._add_buffer("decoder.pos_emb.inv_freq", inv_freq)
        self._sl = self._slices()

    # ---- helpers ---------------------------------------------------------------------------
    def _ref_order_key(self, name, index):
        # RelPartialLearnableMultiHeadAttn registers qkv_net, o_net, layer_norm (base class) and then r_net
        # (optimus_txl_decoder.py:223-240,305-310); here r_net sits next to qkv_net in the flat buffer
        if ".dec_attn.r_net." in name:
            return index + 3.5
        return index

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
        return float(self.dec_dropout) if self.training else 0.0

    def _pos_table(self, klen, dev):
        inv_freq = self.decoder.pos_emb.inv_freq.to(dev)
        pos_seq = torch.arange(klen - 1, -1, -1.0, device=dev, dtype=torch.float32)
        s = torch.outer(pos_seq, inv_freq)
        return torch.cat([s.sin(), s.cos()], dim=-1)          # [klen, d] fp32 constant table (host-side plumbing)

    @staticmethod
    def _mems_in(dec_mems):
        """reference calling conventions: tuple() / None / (list,) / list of n_layer+1 tensors"""
        if dec_mems is None:
            return None
        if isinstance(dec_mems, tuple) and len(dec_mems) == 1 and isinstance(dec_mems[0], (list, tuple)):
            dec_mems = dec_mems[0]
        dec_mems = list(dec_mems)
        return dec_mems if len(dec_mems) > 0 else None

    # ---- reference-facing API ----------------------------------------------------------------
    def layer_grad_range(self, l):
        return self.grad_range("decoder.layers.%d." % l)

    def forward(self, dec_input, dec_mems, dec_seg_len=None, return_avg_attn=False):
        if return_avg_attn:
            raise NotImplementedError("return_avg_attn materialises the T x T attention matrix; not on the hot path")
        if not dec_input.is_cuda:
            raise RuntimeError("emo_disentanger_b200 models run on CUDA only (no CPU fallback); call .cuda()")
        mems = self._mems_in(dec_mems)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            anchor = dict(self.named_parameters())["dec_out_proj.bias"]
            logits, new_mems = ModelFn.apply(anchor, self, (dec_input, mems)), None
            new_mems = self._last_new_mems
        else:
            (logits, new_mems), _ = self._forward_impl(dec_input, mems, save=False)
        return logits, new_mems

    def generate(self, dec_input, dec_mems):
        """dec_input [t, 1] -> (logits [V] of the last position, new_mems)   (plain_transformer.py:51-58)"""
        with torch.no_grad():
            (logits, new_mems), _ = self._forward_impl(dec_input, self._mems_in(dec_mems), save=False, last_only=True)
        return logits, new_mems

    def compute_loss(self, dec_logits, dec_tgt, reduction='mean'):
        if reduction != 'mean':
            raise NotImplementedError("only reduction='mean' is used by the reference")
        ce_loss = CrossEntropyFn.apply(dec_logits, dec_tgt, self.pad_index, False)
        return {'ce_loss': ce_loss, 'total_loss': ce_loss}

    def train_step(self, dec_input, dec_tgt, gscale=1.0, count_allreduce=None):
        """fused fwd + CE + bwd (no autograd); dec_input / dec_tgt [T, B]; returns [n_valid, loss_sum, n_correct]."""
        self._prepare_grads()
        hid, saved = self._forward_hidden(dec_input, None, save=True)
        logits = self._logits(hid)
        acc = torch.zeros(3, dtype=torch.float32, device=dec_input.device)
        ops.ce_count(dec_tgt, self.pad_index, acc[0:1], batch_first=False)
        count = acc[0:1]
        if count_allreduce is not None:
            count = count_allreduce(acc[0:1])
        dl = torch.empty(hid.shape[0], self.ldv, dtype=self.compute_dtype, device=hid.device)
        ops.ce_fwd_bwd(logits, dec_tgt, self.vocab_size, self.pad_index, count, acc[1:2], acc[2:3], None, dl, gscale,
                       batch_first=False)
        self._backward_from_dl(saved, hid, dl)
        return acc

    # ---- forward -------------------------------------------------------------------------------
    def _logits(self, hid):
        Wc = self.weights()
        logits = torch.empty(hid.shape[0], self.ldv, dtype=torch.float32, device=hid.device)
        ops.linear_fwd(hid, self._wv(Wc, "dec_out_proj.weight"), logits[:, :self.vocab_size],
                       bias=self._wv(self._flat, "dec_out_proj.bias"))
        return logits

    def _forward_impl(self, tok, mems, save, last_only=False):
        T, B = tok.shape
        hid, saved = self._forward_hidden(tok, mems, save)
        new_mems = self._last_new_mems
        if last_only:
            row = hid.view(B, T, -1)[0, T - 1:T]                       # [-1, 0, :] of the [T,B,V] logits
            return (self._logits(row.contiguous())[0, :self.vocab_size], new_mems), saved
        logits = self._logits(hid)
        out = logits[:, :self.vocab_size].view(B, T, self.vocab_size).transpose(0, 1)     # [T, B, V] view
        if save:
            saved["hid"] = hid
        return out if save else (out, new_mems), saved

    def _forward_hidden(self, tok, mems, save):
        T, B = tok.shape
        d, f, H, L = self.dec_d_model, self.dec_d_ff, self.dec_n_head, self.dec_n_layer
        dt, dev = self.compute_dtype, tok.device
        Wc, Wf = self.weights(), self._flat
        p = self._p_drop()
        seed = self.next_seed()
        have_mems = self.dec_mem_len > 0
        mlen = mems[0].shape[0] if (mems is not None and mems[0].dim() == 3) else 0
        if save and mlen > 0:
            raise NotImplementedError("training with a non-empty memory is not used by the reference (mem_len: 0)")
        klen = mlen + T
        new = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
        # embedding: drop(drop(E[tok] * sqrt(d)))   (plain_transformer.py:62 and optimus_txl_decoder.py:798)
        h = new(B * T, d)
        ops.embed_fwd(tok, None, self._wv(Wf, "word_emb.emb_lookup.weight"), None, None, h, d ** 0.5, p,
                      site_seed(seed, 0), batch_first=False)
        if p > 0:
            ops.dropout_apply(h, h, p, site_seed(seed, 1))
        pos = self._pos_table(klen, dev).to(dt)
        if p > 0:
            ops.dropout_apply(pos, pos, p, site_seed(seed, 2))
        rw, rr = self._wv(Wf, "decoder.r_w_bias"), self._wv(Wf, "decoder.r_r_bias")
        hids = [h]
        layers = []
        scale = 1.0 / (E ** 0.5)
        for l in range(L):
            nm = "decoder.layers.%d." % l
            if mlen > 0:
                mem = mems[l].transpose(0, 1).to(dt)                                    # [B, mlen, d]
                cat = torch.cat([mem, h.view(B, T, d)], dim=1).reshape(B * klen, d)     # layout plumbing
            else:
                cat = h
            a, m1, r1 = new(B * klen, d), new(B * klen, dtype=torch.float32), new(B * klen, dtype=torch.float32)
            ops.ln_fwd(cat, self._wv(Wf, nm + "dec_attn.layer_norm.weight"), self._wv(Wf, nm + "dec_attn.layer_norm.bias"), a, m1, r1)
            heads = new(B * klen, 3 * d)
            ops.linear_fwd(a, self._wv(Wc, nm + "dec_attn.qkv_net.weight"), heads)
            h3 = heads.view(B, klen, 3 * d)
            q = h3[:, mlen:, 0:d].unflatten(-1, (H, E))
            if mlen > 0 and B > 1:
                q = q.contiguous()
            k = h3[:, :, d:2 * d].unflatten(-1, (H, E))
            v = h3[:, :, 2 * d:3 * d].unflatten(-1, (H, E))
            r = new(klen, d)
            ops.linear_fwd(pos, self._wv(Wc, nm + "dec_attn.r_net.weight"), r)
            att = new(B * T, d)
            lse = new(B, H, T, dtype=torch.float32) if save else None
            ops.relattn_fwd(q, k, v, r.view(klen, H, E), rw, rr, att.view(B, T, d), lse, scale, p, site_seed(seed, 4 + 4 * l))
            h1 = new(B * T, d)
            ops.linear_fwd(att, self._wv(Wc, nm + "dec_attn.o_net.weight"), h1, drop_p=p, seed=site_seed(seed, 5 + 4 * l),
                           residual=h, ld_res=d)
            c, m2, r2 = new(B * T, d), new(B * T, dtype=torch.float32), new(B * T, dtype=torch.float32)
            ops.ln_fwd(h1, self._wv(Wf, nm + "pos_ff.layer_norm.weight"), self._wv(Wf, nm + "pos_ff.layer_norm.bias"), c, m2, r2)
            ff = new(B * T, f)
            ops.linear_fwd(c, self._wv(Wc, nm + "pos_ff.CoreNet.0.weight"), ff, bias=self._wv(Wf, nm + "pos_ff.CoreNet.0.bias"),
                           act=ops.ACT_RELU, drop_p=p, seed=site_seed(seed, 6 + 4 * l))
            h2 = new(B * T, d)
            ops.linear_fwd(ff, self._wv(Wc, nm + "pos_ff.CoreNet.3.weight"), h2, bias=self._wv(Wf, nm + "pos_ff.CoreNet.3.bias"),
                           drop_p=p, seed=site_seed(seed, 7 + 4 * l), residual=h1, ld_res=d)
            if save:
                layers.append((h, m1, r1, a, heads, r, att, lse, h1, m2, r2, c, ff))
            h = h2
            hids.append(h)
        # memory update (optimus_txl_decoder.py:702-722): last mem_len rows of cat[mem, hid] per layer input
        self._last_new_mems = None
        if have_mems:
            end, beg = klen, max(0, klen - self.dec_mem_len)
            nm_ = []
            for i in range(L + 1):
                hv = hids[i].view(B, T, d)
                full = torch.cat([mems[i].transpose(0, 1).to(dt), hv], dim=1) if mlen > 0 else hv
                nm_.append(full[:, beg:end].detach().contiguous().transpose(0, 1))        # exposed as [m, B, d]
            self._last_new_mems = nm_
        out = h
        if p > 0:
            out = ops.dropout_apply(h, new(B * T, d), p, site_seed(seed, 3))
        saved = None
        if save:
            saved = {"layers": layers, "tokens": tok, "seed": seed, "p": p, "B": B, "T": T, "pos": pos}
        return out, saved

    # ---- backward ------------------------------------------------------------------------------
    def _backward_impl(self, saved, dlogits):
        """autograd path: dlogits [T,B,V] fp32 (any strides) -> padded batch-major compute-dtype buffer."""
        T, B, V = dlogits.shape
        dl = torch.zeros(B * T, self.ldv, dtype=self.compute_dtype, device=dlogits.device)
        dl.view(B, T, self.ldv)[:, :, :V].copy_(dlogits.transpose(0, 1))
        self._backward_from_dl(saved, saved["hid"], dl)

    def _backward_from_dl(self, saved, hid, dl):
        B, T = saved["B"], saved["T"]
        d, f, H, L, V = self.dec_d_model, self.dec_d_ff, self.dec_n_head, self.dec_n_layer, self.vocab_size
        dt, dev = self.compute_dtype, dl.device
        Wc, Wf = self.weights(), self._flat
        p, seed, pos = saved["p"], saved["seed"], saved["pos"]
        R = B * T
        new = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
        keep_scale = 1.0 / (1.0 - p)
        scale = 1.0 / (E ** 0.5)
        dlv = dl[:, :V]
        ops.linear_wgrad(dlv, hid, self._gv("dec_out_proj.weight"))
        ops.colsum(dl, self._gv("dec_out_proj.bias"), n=V)
        dh = new(R, d)
        ops.linear_dgrad(dlv, self._wv(Wc, "dec_out_proj.weight"), dh)
        if p > 0:
            ops.dropout_apply(dh, dh, p, site_seed(seed, 3))
        rw, rr = self._wv(Wf, "decoder.r_w_bias"), self._wv(Wf, "decoder.r_r_bias")
        for l in reversed(range(L)):
            nm = "decoder.layers.%d." % l
            h, m1, r1, a, heads, r, att, lse, h1, m2, r2, c, ff = saved["layers"][l]
            # h2 = h1 + drop(ff W2^T + b2)
            g2 = dh
            if p > 0:
                g2 = ops.dropout_apply(dh, new(R, d), p, site_seed(seed, 7 + 4 * l))
            ops.colsum(g2, self._gv(nm + "pos_ff.CoreNet.3.bias"))
            ops.linear_wgrad(g2, ff, self._gv(nm + "pos_ff.CoreNet.3.weight"))
            dff = new(R, f)
            ops.linear_dgrad(g2, self._wv(Wc, nm + "pos_ff.CoreNet.3.weight"), dff, act=ops.ACT_RELU_MASK_BWD, aux=ff,
                             ld_aux=f, aux_scale=keep_scale, colsum_out=self._gv(nm + "pos_ff.CoreNet.0.bias"))
            ops.linear_wgrad(dff, c, self._gv(nm + "pos_ff.CoreNet.0.weight"))
            dc = new(R, d)
            ops.linear_dgrad(dff, self._wv(Wc, nm + "pos_ff.CoreNet.0.weight"), dc)
            dh1 = new(R, d)
            dh1d = new(R, d) if p > 0 else None
            ops.ln_bwd(dc, h1, m2, r2, self._wv(Wf, nm + "pos_ff.layer_norm.weight"), dh1, self._gv(nm + "pos_ff.layer_norm.weight"),
                       self._gv(nm + "pos_ff.layer_norm.bias"), add_in=dh, dx_drop=dh1d, drop_p=p, seed=site_seed(seed, 5 + 4 * l))
            go = dh1d if p > 0 else dh1
            # h1 = h + drop(att Wo^T)
            ops.linear_wgrad(go, att, self._gv(nm + "dec_attn.o_net.weight"))
            datt = dc
            ops.linear_dgrad(go, self._wv(Wc, nm + "dec_attn.o_net.weight"), datt)
            dheads = new(R, 3 * d)
            h3, dh3 = heads.view(B, T, 3 * d), dheads.view(B, T, 3 * d)
            q, k, v = (h3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            dq, dk, dv = (dh3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            dr = torch.zeros(T, H, E, dtype=torch.float32, device=dev)
            ops.relattn_bwd(q, k, v, r.view(T, H, E), rw, rr, att.view(B, T, d), datt.view(B, T, d), lse, dq, dk, dv, dr,
                            self._gv("decoder.r_w_bias"), self._gv("decoder.r_r_bias"), scale, p, site_seed(seed, 4 + 4 * l))
            drc = dr.view(T, d) if dt == torch.float32 else ops.cast(dr.view(T, d), new(T, d))
            ops.linear_wgrad(drc, pos, self._gv(nm + "dec_attn.r_net.weight"))
            ops.linear_wgrad(dheads, a, self._gv(nm + "dec_attn.qkv_net.weight"))
            da = new(R, d)
            ops.linear_dgrad(dheads, self._wv(Wc, nm + "dec_attn.qkv_net.weight"), da)
            dx = new(R, d)
            ops.ln_bwd(da, h, m1, r1, self._wv(Wf, nm + "dec_attn.layer_norm.weight"), dx, self._gv(nm + "dec_attn.layer_norm.weight"),
                       self._gv(nm + "dec_attn.layer_norm.bias"), add_in=dh1)
            dh = dx
            self._layer_done(l)
        if p > 0:
            ops.dropout_apply(dh, dh, p, site_seed(seed, 1))
        ops.embed_bwd(saved["tokens"], None, dh, self._gv("word_emb.emb_lookup.weight"), None, d ** 0.5, p,
                      site_seed(seed, 0), pad_idx=self.pad_index, batch_first=False)
