"""MusicGPT2 -- drop-in for reference stage2_accompaniment/model/music_gpt2.py (12 x HF GPT2Block,
no wpe, no ln_f), computing on hand-written sm_100a kernels through libemo_b200.so.

Same constructor / forward / compute_loss signatures and the same state-dict keys as the reference
(`transformer_decoder.{l}.ln_1.weight`, `.attn.c_attn.weight [512,1536]` in HF Conv1D [in,out] layout,
`.attn.c_proj`, `.mlp.c_fc`, `.mlp.c_proj`, ...).  Checkpoints written under transformers 4.28 also
carry the per-layer causal-mask buffers `attn.bias` / `attn.masked_bias`; they are dropped on load.

Per layer (pre-LN GPT2Block):
  a    = LN1(x)
  qkv  = a Wattn + b                       tcgen05 GEMM (Conv1D layout -> NN contraction)
  att  = softmax(q k^T / 8 + causal) v     attn.cu (flash style; attention-prob dropout in-kernel)
  h    = x + drop(att Wproj + b)           GEMM epilogue: bias + dropout + residual
  c    = LN2(h)
  g    = gelu_new(c Wfc + b)               GEMM epilogue: bias + gelu_new (pre-activation kept for bwd)
  out  = h + drop(g Wproj2 + b)            GEMM epilogue
"""
import torch

from .. import ops
from ..engine import site_seed
from .base import Stage2Base, _normal, _zeros

E = 64


class MusicGPT2(Stage2Base):
    def __init__(self, n_token, n_layer, n_head, d_model, d_ff, d_embed,
                 activation='relu', dropout=0.1, use_pe=True,
                 use_segment_emb=False, n_segment_types=None, use_chord_mhot_emb=False,
                 compute_dtype=torch.bfloat16):
        super().__init__(n_token, n_layer, n_head, d_model, d_ff, d_embed, activation, dropout, use_pe,
                         use_segment_emb, n_segment_types, use_chord_mhot_emb, compute_dtype)
        d, f = d_model, d_ff
        for l in range(n_layer):
            p = "transformer_decoder.%d." % l
            self._add_param(p + "ln_1.weight", (d,), _normal(0.01, 1.0))
            self._add_param(p + "ln_1.bias", (d,), _zeros)
            self._add_param(p + "attn.c_attn.weight", (d, 3 * d), _normal(0.02))     # HF Conv1D: [in, out]
            self._add_param(p + "attn.c_attn.bias", (3 * d,), _zeros)
            self._add_param(p + "attn.c_proj.weight", (d, d), _normal(0.02))
            self._add_param(p + "attn.c_proj.bias", (d,), _zeros)
            self._add_param(p + "ln_2.weight", (d,), _normal(0.01, 1.0))
            self._add_param(p + "ln_2.bias", (d,), _zeros)
            self._add_param(p + "mlp.c_fc.weight", (d, f), _normal(0.02))
            self._add_param(p + "mlp.c_fc.bias", (f,), _zeros)
            self._add_param(p + "mlp.c_proj.weight", (f, d), _normal(0.02))
            self._add_param(p + "mlp.c_proj.bias", (d,), _zeros)
        self._finish()

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {k: v for k, v in state_dict.items()
              if not (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))}
        return super().load_state_dict(sd, strict=strict, **kw)

    # ---- forward -------------------------------------------------------------------------------
    def _forward_hidden(self, x, seg, save):
        B, T = x.shape
        R, d, f, H = B * T, self.d_model, self.d_ff, self.n_head
        dt, dev = self.compute_dtype, x.device
        Wc, Wf = self.weights(), self._flat
        p = self._p_drop()
        seed = self.next_seed()
        h_in = self._embed(x, seg, seed)
        layers = []
        new = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
        scale = 1.0 / (E ** 0.5)
        for l in range(self.n_layer):
            nm = "transformer_decoder.%d." % l
            a, m1, r1 = new(R, d), new(R, dtype=torch.float32), new(R, dtype=torch.float32)
            ops.ln_fwd(h_in, self._wv(Wf, nm + "ln_1.weight"), self._wv(Wf, nm + "ln_1.bias"), a, m1, r1)
            qkv = new(R, 3 * d)
            ops.linear_fwd_t(a, self._wv(Wc, nm + "attn.c_attn.weight"), qkv, bias=self._wv(Wf, nm + "attn.c_attn.bias"))
            q3 = qkv.view(B, T, 3 * d)
            q, k, v = (q3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            att = new(R, d)
            lse = new(B, H, T, dtype=torch.float32)
            ops.attn_fwd(q, k, v, att.view(B, T, d), lse, scale, p, site_seed(seed, 4 * l + 1))
            h = new(R, d)
            ops.linear_fwd_t(att, self._wv(Wc, nm + "attn.c_proj.weight"), h, bias=self._wv(Wf, nm + "attn.c_proj.bias"),
                             drop_p=p, seed=site_seed(seed, 4 * l + 2), residual=h_in, ld_res=d)
            c, m2, r2 = new(R, d), new(R, dtype=torch.float32), new(R, dtype=torch.float32)
            ops.ln_fwd(h, self._wv(Wf, nm + "ln_2.weight"), self._wv(Wf, nm + "ln_2.bias"), c, m2, r2)
            u = new(R, f) if save else None
            g = new(R, f)
            ops.linear_fwd_t(c, self._wv(Wc, nm + "mlp.c_fc.weight"), g, bias=self._wv(Wf, nm + "mlp.c_fc.bias"),
                             act=ops.ACT_GELU_NEW, aux_out=u, ld_aux=f)
            out = new(R, d)
            ops.linear_fwd_t(g, self._wv(Wc, nm + "mlp.c_proj.weight"), out, bias=self._wv(Wf, nm + "mlp.c_proj.bias"),
                             drop_p=p, seed=site_seed(seed, 4 * l + 3), residual=h, ld_res=d)
            if save:
                layers.append((h_in, m1, r1, a, qkv, att, lse, h, m2, r2, c, u, g))
            h_in = out
        saved = None
        if save:
            saved = {"layers": layers, "tokens": x, "seg": seg, "seed": seed, "p": p, "B": B, "T": T}
        return h_in, saved

    # ---- backward ------------------------------------------------------------------------------
    def _backward_hidden(self, saved, dout):
        B, T = saved["B"], saved["T"]
        R, d, f, H = B * T, self.d_model, self.d_ff, self.n_head
        dt, dev = self.compute_dtype, dout.device
        Wc, Wf = self.weights(), self._flat
        p, seed = saved["p"], saved["seed"]
        new = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
        scale = 1.0 / (E ** 0.5)
        for l in reversed(range(self.n_layer)):
            nm = "transformer_decoder.%d." % l
            h_in, m1, r1, a, qkv, att, lse, h, m2, r2, c, u, g = saved["layers"][l]
            # out = h + drop(g Wproj2 + b)
            gp2 = dout
            if p > 0:
                gp2 = ops.dropout_apply(dout, new(R, d), p, site_seed(seed, 4 * l + 3))
            ops.colsum(gp2, self._gv(nm + "mlp.c_proj.bias"))
            ops.linear_wgrad_t(gp2, g, self._gv(nm + "mlp.c_proj.weight"))
            du = new(R, f)
            ops.linear_dgrad_t(gp2, self._wv(Wc, nm + "mlp.c_proj.weight"), du, act=ops.ACT_GELU_NEW_BWD, aux=u, ld_aux=f,
                               colsum_out=self._gv(nm + "mlp.c_fc.bias"))            # bias gradient rides in the epilogue
            ops.linear_wgrad_t(du, c, self._gv(nm + "mlp.c_fc.weight"))
            dc = new(R, d)
            ops.linear_dgrad_t(du, self._wv(Wc, nm + "mlp.c_fc.weight"), dc)
            # c = LN2(h); h also feeds the residual -> dh = LN2'(dc) + dout
            dh = new(R, d)
            dhd = new(R, d) if p > 0 else None
            ops.ln_bwd(dc, h, m2, r2, self._wv(Wf, nm + "ln_2.weight"), dh, self._gv(nm + "ln_2.weight"),
                       self._gv(nm + "ln_2.bias"), add_in=dout, dx_drop=dhd, drop_p=p, seed=site_seed(seed, 4 * l + 2),
                       dxsum=self._gv(nm + "attn.c_proj.bias"))
            gp = dhd if p > 0 else dh
            ops.linear_wgrad_t(gp, att, self._gv(nm + "attn.c_proj.weight"))
            datt = dc      # reuse
            ops.linear_dgrad_t(gp, self._wv(Wc, nm + "attn.c_proj.weight"), datt)
            dqkv = new(R, 3 * d)
            q3, dq3 = qkv.view(B, T, 3 * d), dqkv.view(B, T, 3 * d)
            q, k, v = (q3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            dq, dk, dv = (dq3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            ops.attn_bwd(q, k, v, att.view(B, T, d), datt.view(B, T, d), lse, dq, dk, dv, scale, p,
                         site_seed(seed, 4 * l + 1))
            ops.colsum(dqkv, self._gv(nm + "attn.c_attn.bias"))
            ops.linear_wgrad_t(dqkv, a, self._gv(nm + "attn.c_attn.weight"))
            da = new(R, d)
            ops.linear_dgrad_t(dqkv, self._wv(Wc, nm + "attn.c_attn.weight"), da)
            dx = new(R, d)
            ops.ln_bwd(da, h_in, m1, r1, self._wv(Wf, nm + "ln_1.weight"), dx, self._gv(nm + "ln_1.weight"),
                       self._gv(nm + "ln_1.bias"), add_in=dh)
            dout = dx
            self._layer_done(l)
        return dout

    def layer_grad_range(self, l):
        return self.grad_range("transformer_decoder.%d." % l)
