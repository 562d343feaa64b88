"""MusicPerformer -- drop-in for reference stage2_accompaniment/model/music_performer.py
(+ fast_transformer_decoder.py and the fast_transformers layers it builds), computing on
hand-written sm_100a kernels through libemo_b200.so.

Same constructor / forward / compute_loss signatures and the same state-dict keys
(`token_emb.emb_lookup.weight`, `transformer_decoder.decoder_layers.{l}.attention.query_projection.weight`,
..., `...inner_attention.feature_map.omega`), so reference checkpoints load unchanged.

Per layer (post-LN TransformerEncoderLayer):
  qkv  = x Wqkv^T + b                       tcgen05 GEMM (q,k,v projections packed: one launch)
  att  = FAVOR+ causal linear attention     favor.cu (phi recomputed in-kernel)
  s1   = x + drop(att Wo^T + bo)            GEMM epilogue: bias + dropout + residual
  y1   = LN1(s1)
  h    = drop(relu(y1 W1^T + b1))           GEMM epilogue: bias + relu + dropout
  s2   = y1 + drop(h W2^T + b2)             GEMM epilogue
  out  = LN2(s2)
"""
import os

import torch

from .. import ops
from ..engine import site_seed
from .base import Stage2Base, _normal, _zeros

FUSE_LN = os.environ.get("EMO_FUSE_LN", "1") != "0"

E = 64  # head dim


class MusicPerformer(Stage2Base):
    def __init__(self, n_token, n_layer, n_head, d_model, d_ff, d_embed,
                 activation='relu', dropout=0.1, use_pe=True, favor_feature_dims=None,
                 use_segment_emb=False, n_segment_types=None, use_chord_mhot_emb=False,
                 compute_dtype=torch.bfloat16, orthogonal_features=False):
        super().__init__(n_token, n_layer, n_head, d_model, d_ff, d_embed, activation, dropout, use_pe,
                         use_segment_emb, n_segment_types, use_chord_mhot_emb, compute_dtype)
        self.favor_feature_dims = 2 * d_model // n_head if favor_feature_dims is None else favor_feature_dims
        if self.favor_feature_dims != 128:
            raise ValueError("favor_feature_dims must be 128 (reference YAML feature_map.n_dims)")
        if activation != 'relu':
            raise NotImplementedError("the reference only builds the Performer with activation='relu'")
        self.orthogonal_features = orthogonal_features
        d, f = d_model, d_ff
        for l in range(n_layer):
            p = "transformer_decoder.decoder_layers.%d." % l
            # q, k, v weights (and biases) are adjacent in the flat buffer -> one [1536, 512] GEMM
            for nm in ("query", "key", "value"):
                self._add_param(p + "attention.%s_projection.weight" % nm, (d, d), _normal(0.01))
            for nm in ("query", "key", "value"):
                self._add_param(p + "attention.%s_projection.bias" % nm, (d,), _zeros)
            self._add_param(p + "attention.out_projection.weight", (d, d), _normal(0.01))
            self._add_param(p + "attention.out_projection.bias", (d,), _zeros)
            self._add_param(p + "linear1.weight", (f, d), _normal(0.01))
            self._add_param(p + "linear1.bias", (f,), _zeros)
            self._add_param(p + "linear2.weight", (d, f), _normal(0.01))
            self._add_param(p + "linear2.bias", (d,), _zeros)
            self._add_param(p + "norm1.weight", (d,), _normal(0.01, 1.0))
            self._add_param(p + "norm1.bias", (d,), _zeros)
            self._add_param(p + "norm2.weight", (d,), _normal(0.01, 1.0))
            self._add_param(p + "norm2.bias", (d,), _zeros)
        self._finish()
        # omega buffers: views of one [L, 64, 64] tensor (state-dict keys kept; ignored on load by
        # the reference scripts, train.py:306-308)
        self._omegas = torch.zeros(n_layer, E, self.favor_feature_dims // 2)
        for l in range(n_layer):
            self._add_buffer("transformer_decoder.decoder_layers.%d.attention.inner_attention.feature_map.omega" % l,
                             self._omegas[l])
        self.fixed_omegas = None      # test / decode hook: use these instead of redrawing

    # ---- feature map draw (Favor.new_feature_map; redrawn on EVERY forward like the reference) --
    def draw_omegas(self, device):
        if self.fixed_omegas is not None:
            return self.fixed_omegas.to(device=device, dtype=torch.float32).contiguous()
        L = self.n_layer
        g = torch.randn(L, E, E, device=device, dtype=torch.float32)
        if not self.orthogonal_features:
            return g
        q, _ = torch.linalg.qr(g)
        return (q * g.norm(dim=2)[:, None, :]).contiguous()

    def _layer_names(self, l):
        return "transformer_decoder.decoder_layers.%d." % l

    def _qkv_w(self, buf, l):
        cache = self.__dict__.setdefault("_view_cache", {})
        key = (buf.data_ptr(), "qkv_w", l)
        v = cache.get(key)
        if v is None:
            off, n, _ = self._sl[self._layer_names(l) + "attention.query_projection.weight"]
            v = cache[key] = buf[off:off + 3 * n].view(3 * self.d_model, self.d_model)
        return v

    def _qkv_b(self, buf, l):
        cache = self.__dict__.setdefault("_view_cache", {})
        key = (buf.data_ptr(), "qkv_b", l)
        v = cache.get(key)
        if v is None:
            off, n, _ = self._sl[self._layer_names(l) + "attention.query_projection.bias"]
            v = cache[key] = buf[off:off + 3 * n]
        return v

    # ---- forward -------------------------------------------------------------------------------
    def _forward_hidden(self, x, seg, save):
        B, T = x.shape
        R, d, f, H = B * T, self.d_model, self.d_ff, self.n_head
        dt, dev = self.compute_dtype, x.device
        Wc, Wf = self.weights(), self._flat
        p = self._p_drop()
        seed = self.next_seed()
        omegas = self.draw_omegas(dev)
        self.last_omegas = omegas
        h_in = self._embed(x, seg, seed)
        layers = []
        new = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
        # EMO_FUSE_LN=0: out-projection / linear2 and their residual + LayerNorm as two launches each (A/B switch)
        fuse_ln = FUSE_LN and dt == torch.bfloat16 and d == 512 and R > 128
        for l in range(self.n_layer):
            nm = self._layer_names(l)
            qkv = new(R, 3 * d)
            ops.linear_fwd(h_in, self._qkv_w(Wc, l), qkv, bias=self._qkv_b(Wf, l))
            q3 = qkv.view(B, T, 3 * d)
            q, k, v = (q3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            att = new(R, d)
            den = new(B, T, H, dtype=torch.float32) if save else None
            state = ops.favor_workspace(B, T, H, dt, dev)      # segment-state sums (kept for backward when save)
            ops.favor_fwd(q, k, v, omegas[l], att.view(B, T, d), den, seg_states=state)
            # the residual joins inside the LN kernel, in fp32: x + dropout(attn) is never rounded to bf16 on the
            # way into norm1 / norm2 (s1 / s2 below are the copies the backward re-normalises)
            y1, m1, r1 = new(R, d), new(R, dtype=torch.float32), new(R, dtype=torch.float32)
            if fuse_ln:
                # projection + dropout + residual + LayerNorm in ONE kernel (emo_gemm_ln_res): whole rows live in tensor
                # memory, the sum is never rounded on its way into the norm and never makes the HBM round trip
                s1 = new(R, d) if save else None
                ops.linear_ln_res_fwd(att, self._wv(Wc, nm + "attention.out_projection.weight"),
                                      self._wv(Wf, nm + "attention.out_projection.bias"), h_in,
                                      self._wv(Wf, nm + "norm1.weight"), self._wv(Wf, nm + "norm1.bias"), y1, m1, r1,
                                      sum_out=s1, drop_p=p, seed=site_seed(seed, 4 * l + 1))
            else:
                b1 = new(R, d)
                ops.linear_fwd(att, self._wv(Wc, nm + "attention.out_projection.weight"), b1,
                               bias=self._wv(Wf, nm + "attention.out_projection.bias"),
                               drop_p=p, seed=site_seed(seed, 4 * l + 1))
                s1 = b1 if save else None            # in place: each lane rewrites exactly the elements it has read
                ops.ln_res_fwd(b1, h_in, self._wv(Wf, nm + "norm1.weight"), self._wv(Wf, nm + "norm1.bias"), y1, m1, r1, sum_out=s1)
            hh = new(R, f)
            ops.linear_fwd(y1, self._wv(Wc, nm + "linear1.weight"), hh, bias=self._wv(Wf, nm + "linear1.bias"),
                           act=ops.ACT_RELU, drop_p=p, seed=site_seed(seed, 4 * l + 2))
            out, m2, r2 = new(R, d), new(R, dtype=torch.float32), new(R, dtype=torch.float32)
            if fuse_ln:
                s2 = new(R, d) if save else None
                ops.linear_ln_res_fwd(hh, self._wv(Wc, nm + "linear2.weight"), self._wv(Wf, nm + "linear2.bias"), y1,
                                      self._wv(Wf, nm + "norm2.weight"), self._wv(Wf, nm + "norm2.bias"), out, m2, r2,
                                      sum_out=s2, drop_p=p, seed=site_seed(seed, 4 * l + 3))
            else:
                b2 = new(R, d)
                ops.linear_fwd(hh, self._wv(Wc, nm + "linear2.weight"), b2, bias=self._wv(Wf, nm + "linear2.bias"),
                               drop_p=p, seed=site_seed(seed, 4 * l + 3))
                s2 = b2 if save else None
                ops.ln_res_fwd(b2, y1, self._wv(Wf, nm + "norm2.weight"), self._wv(Wf, nm + "norm2.bias"), out, m2, r2, sum_out=s2)
            if save:
                layers.append((h_in, qkv, att, den, state, s1, m1, r1, y1, hh, s2, m2, r2))
            h_in = out
        saved = None
        if save:
            saved = {"layers": layers, "omegas": omegas, "tokens": x, "seg": seg, "seed": seed, "p": p, "B": B, "T": T}
        return h_in, saved

    # ---- backward ------------------------------------------------------------------------------
    def _backward_hidden(self, saved, dout):
        B, T = saved["B"], saved["T"]
        R, d, f, H = B * T, self.d_model, self.d_ff, self.n_head
        dt, dev = self.compute_dtype, dout.device
        Wc, Wf = self.weights(), self._flat
        p, seed, omegas = saved["p"], saved["seed"], saved["omegas"]
        new = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
        keep_scale = 1.0 / (1.0 - p)
        for l in reversed(range(self.n_layer)):
            nm = self._layer_names(l)
            h_in, qkv, att, den, state, s1, m1, r1, y1, hh, s2, m2, r2 = saved["layers"][l]
            # out = LN2(s2), s2 = y1 + drop(h W2^T + b2)
            ds2 = new(R, d)
            ds2d = new(R, d) if p > 0 else None
            ops.ln_bwd(dout, s2, m2, r2, self._wv(Wf, nm + "norm2.weight"), ds2, self._gv(nm + "norm2.weight"),
                       self._gv(nm + "norm2.bias"), dx_drop=ds2d, drop_p=p, seed=site_seed(seed, 4 * l + 3),
                       dxsum=self._gv(nm + "linear2.bias"))                  # bias gradients ride along
            g2 = ds2d if p > 0 else ds2
            ops.linear_wgrad(g2, hh, self._gv(nm + "linear2.weight"))
            da = new(R, f)
            ops.linear_dgrad(g2, self._wv(Wc, nm + "linear2.weight"), da, act=ops.ACT_RELU_MASK_BWD, aux=hh,
                             ld_aux=f, aux_scale=keep_scale, colsum_out=self._gv(nm + "linear1.bias"))
            ops.linear_wgrad(da, y1, self._gv(nm + "linear1.weight"))
            dy1 = new(R, d)
            ops.linear_dgrad(da, self._wv(Wc, nm + "linear1.weight"), dy1, residual=ds2, ld_res=d)
            # y1 = LN1(s1), s1 = h_in + drop(att Wo^T + bo)
            ds1 = ds2      # reuse buffers
            ds1d = ds2d
            ops.ln_bwd(dy1, s1, m1, r1, self._wv(Wf, nm + "norm1.weight"), ds1, self._gv(nm + "norm1.weight"),
                       self._gv(nm + "norm1.bias"), dx_drop=ds1d, drop_p=p, seed=site_seed(seed, 4 * l + 1),
                       dxsum=self._gv(nm + "attention.out_projection.bias"))
            g1 = ds1d if p > 0 else ds1
            ops.linear_wgrad(g1, att, self._gv(nm + "attention.out_projection.weight"))
            datt = dy1     # reuse
            ops.linear_dgrad(g1, self._wv(Wc, nm + "attention.out_projection.weight"), datt)
            dqkv = new(R, 3 * d)
            q3, dq3 = qkv.view(B, T, 3 * d), dqkv.view(B, T, 3 * d)
            q, k, v = (q3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            dq, dk, dv = (dq3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            ops.favor_bwd(q, k, v, omegas[l], att.view(B, T, d), datt.view(B, T, d), den, state, dq, dk, dv)
            ops.colsum(dqkv, self._qkv_b(self._flat_grad, l))
            ops.linear_wgrad(dqkv, h_in, self._qkv_w(self._flat_grad, l))
            dx = new(R, d)
            ops.linear_dgrad(dqkv, self._qkv_w(Wc, l), dx, residual=ds1, ld_res=d)
            dout = dx
            self._layer_done(l)
        return dout

    def layer_grad_range(self, l):
        return self.grad_range(self._layer_names(l))
