"""Incremental autoregressive decode for the stage-2 models.

The reference re-runs the model over the WHOLE prefix for every sampled token
(stage2_accompaniment/inference.py:252-272: Performer O(t), GPT-2 O(t^2) per token) and copies the last
logits row to the host.  Here the prefix is folded into per-layer state:

  * Performer: the FAVOR+ prefix state  [sum phi(k) v^T | sum phi(k)]  (128 x 80 fp32 per head), advanced
    by favor.cu's recurrent step (one token) or chunked forward with state_in/state_out (a block of tokens:
    the primer, or a lead-sheet bar appended mid-generation, inference.py:293-307).  A step for a RAGGED
    batch of independent sequences is one static-shaped launch sequence -> captured once in a CUDA graph.
  * GPT-2: a K/V cache per layer; attention over the cache for the new rows only.

The feature map Omega is drawn ONCE per decoder (the reference passes `omit_feature_map_draw: steps > 0`
to keep it fixed during a generation, but fast_transformer_decoder.py:62,71 drops the kwarg and Omega is
redrawn on every call; a recurrent state is only meaningful with the intended, fixed, Omega).
Validity: absolute positions -> state is valid while len(prefix) <= max_len (inference.py:255-257 slides
the window after that); `generate.py` falls back to full-prefix recompute past that point.
"""
import os

import torch

from . import ops
from .stage2.music_performer import MusicPerformer
from .stage2.music_gpt2 import MusicGPT2

E = 64
# EMO_FUSED_SAMPLER=0: logits projection and sampler as two nodes of the step graph (A/B switch of emo_logits_sample)
FUSED_SAMPLER = os.environ.get("EMO_FUSED_SAMPLER", "1") != "0"


class Stage2Decoder:
    def __init__(self, model, batch=1, max_len=2048, omegas=None, use_graph=True, use_pdl=True, one_kernel=None):
        if model.training:
            raise RuntimeError("decode needs model.eval() (dropout off)")
        self.m = model
        self.B = batch
        self.max_len = max_len
        self.dev = model._flat.device
        self.dt = model.compute_dtype
        self.is_performer = isinstance(model, MusicPerformer)
        if not self.is_performer and not isinstance(model, MusicGPT2):
            raise TypeError("Stage2Decoder drives MusicPerformer or MusicGPT2")
        L, H, d = model.n_layer, model.n_head, model.d_model
        dev = self.dev
        self.pos = torch.zeros(batch, dtype=torch.int64, device=dev)          # next position of each sequence
        self.pos_host = [0] * batch
        if self.is_performer:
            self.omegas = (model.draw_omegas(dev) if omegas is None else omegas.to(dev, torch.float32)).contiguous()
            self.state = torch.zeros(L, batch, H, 128, 80, dtype=torch.float32, device=dev)
        else:
            self.kv = torch.zeros(L, batch, max_len, 2 * d, dtype=self.dt, device=dev)
            self.pos_tok = torch.zeros(batch, dtype=torch.int64, device=dev)   # position of the token a step is processing
            # HF Conv1D stores weights [in, out]; decode keeps [out, in] copies (weights are frozen while
            # generating) so that the one-row-per-sequence step runs the weight-streaming NT kernel
            self.wT = {}
        # static step buffers (graph inputs / outputs)
        self.logits = torch.zeros(batch, model.ldv, dtype=torch.float32, device=dev)
        self.use_graph = bool(use_graph)
        self.use_pdl = bool(use_pdl)      # step kernels overlap their prologue with the predecessor's tail (graph only)
        # the whole step as ONE kernel, a 16-CTA cluster per sequence (bf16 Performer; csrc/decode_step.cu).  Opt-in:
        # bit-identical to the kernel chain, sequences scale almost for free (one cluster each), but at batch 1-4 it
        # measured 312-326 us per step against 200-300 us for the PDL graph of small kernels
        can_one = self.is_performer and self.dt == torch.bfloat16
        self.one_kernel = bool(one_kernel) and can_one
        if self.one_kernel:
            sl = model._sl
            rows = []
            for l in range(L):
                nm = model._layer_names(l)
                rows.append([sl[nm + k][0] for k in (
                    "attention.query_projection.weight", "attention.out_projection.weight", "linear1.weight", "linear2.weight",
                    "attention.query_projection.bias", "attention.out_projection.bias", "linear1.bias", "linear2.bias",
                    "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")])
            self._offs = torch.tensor(rows, dtype=torch.int64, device=dev)
            self._scratch = torch.zeros(batch * 6144, dtype=torch.bfloat16, device=dev)
        self.graph = None
        # pinned staging ring for the step inputs (tokens | segments): a slot is rewritten only after the
        # async H2D copy that read it has completed
        self._ring = [torch.zeros(3, batch, dtype=torch.int64).pin_memory() for _ in range(8)]
        self._ring_ev = [None] * 8
        self._ring_i = 0
        self._dev_in = torch.zeros(3, batch, dtype=torch.int64, device=dev)    # tokens | segments | uniforms (as fp32)
        self.tok_in, self.seg_in = self._dev_in[0], self._dev_in[1]
        self.u_in = self._dev_in[2].view(torch.float32)[:batch]
        # fused step + sample (one graph): sampled ids | status words, read back through pinned memory
        self._sampled = torch.zeros(2 * batch, dtype=torch.int64, device=dev)
        self._sampled_host = torch.zeros(2 * batch, dtype=torch.int64).pin_memory()
        self.sample_cfg = None            # (temperature, top_p, greedy) baked into the fused graph
        self.graph_sample = None

    def _sync_weights(self):
        """Everything that bakes in weight values or addresses (the [out, in] copies of the Conv1D weights, the
        captured step graphs) is rebuilt when the model's masters changed (load_state_dict, an optimizer step) or
        its flat buffer moved (`.to()`); called at the top of every public entry point."""
        m = self.m
        key = (m._weights_version(), m._flat.data_ptr())
        if key == getattr(self, "_wkey", None):
            return
        self._wkey = key
        self.graph = None
        self.graph_sample = None
        if not self.is_performer:
            Wc = m.weights()
            for l in range(m.n_layer):
                for nm in ("attn.c_attn", "attn.c_proj", "mlp.c_fc", "mlp.c_proj"):
                    k = "transformer_decoder.%d.%s.weight" % (l, nm)
                    self.wT[k] = m._wv(Wc, k).t().contiguous()

    def reset(self, b=None):
        sl = slice(None) if b is None else slice(b, b + 1)
        if self.is_performer:
            self.state[:, sl].zero_()
        self.pos[sl] = 0
        for i in (range(self.B) if b is None else [b]):
            self.pos_host[i] = 0

    # ---- shared layer tails -------------------------------------------------------------------
    def _performer_rows(self, h, B, T, favor):
        """h [B*T, d] (embedding rows) -> (s2, norm) of the LAST layer: the hidden state is LN(s2) with `norm`, which
        the caller folds into the logits projection.  `favor(l, qkv, att)` runs the attention.
        Post-LN layers: each LayerNorm is fused into the projection that consumes it (ops.gemm ln=...: in-kernel for
        the <= 8 rows of a decode step, a separate launch for a longer block), so a layer is 5 launches."""
        m = self.m
        d, f, H = m.d_model, m.d_ff, m.n_head
        Wc, Wf = m.weights(), m._flat
        R = B * T
        new = lambda *shape, dtype=self.dt: torch.empty(*shape, dtype=dtype, device=self.dev)
        pre, norm = h, None                 # layer input = LN(pre) with `norm` (None: `pre` is the input itself)
        for l in range(m.n_layer):
            nm = m._layer_names(l)
            qkv = new(R, 3 * d)
            if norm is None:
                h_in = pre
                ops.linear_fwd(h_in, m._qkv_w(Wc, l), qkv, bias=m._qkv_b(Wf, l))
            else:
                h_in = new(R, d)
                ops.linear_fwd(pre, m._qkv_w(Wc, l), qkv, bias=m._qkv_b(Wf, l), ln=(norm[0], norm[1], h_in))
            att = new(R, d)
            favor(l, qkv, att)
            s1 = new(R, d)
            ops.linear_fwd(att, m._wv(Wc, nm + "attention.out_projection.weight"), s1,
                           bias=m._wv(Wf, nm + "attention.out_projection.bias"), residual=h_in, ld_res=d)
            y1 = new(R, d)
            hh = new(R, f)
            ops.linear_fwd(s1, m._wv(Wc, nm + "linear1.weight"), hh, bias=m._wv(Wf, nm + "linear1.bias"), act=ops.ACT_RELU,
                           ln=(m._wv(Wf, nm + "norm1.weight"), m._wv(Wf, nm + "norm1.bias"), y1))
            s2 = new(R, d)
            ops.linear_fwd(hh, m._wv(Wc, nm + "linear2.weight"), s2, bias=m._wv(Wf, nm + "linear2.bias"),
                           residual=y1, ld_res=d)
            pre, norm = s2, (m._wv(Wf, nm + "norm2.weight"), m._wv(Wf, nm + "norm2.bias"))
        return pre, norm

    def _gpt2_rows(self, h, b, T, pos0):
        """one sequence b, T new rows at positions pos0.. -> hidden"""
        m = self.m
        d, f, H = m.d_model, m.d_ff, m.n_head
        Wc, Wf = m.weights(), m._flat
        new = lambda *shape, dtype=self.dt: torch.empty(*shape, dtype=dtype, device=self.dev)
        for l in range(m.n_layer):
            nm = "transformer_decoder.%d." % l
            a = new(T, d)
            ops.ln_fwd(h, m._wv(Wf, nm + "ln_1.weight"), m._wv(Wf, nm + "ln_1.bias"), a)
            qkv = new(T, 3 * d)
            ops.linear_fwd(a, self.wT[nm + "attn.c_attn.weight"], qkv, bias=m._wv(Wf, nm + "attn.c_attn.bias"))
            cache = self.kv[l, b]
            cache[pos0:pos0 + T].copy_(qkv[:, d:])                                   # append K|V rows (plumbing)
            Tk = pos0 + T
            q = qkv[:, :d].view(1, T, H, E)
            k = cache[:Tk, :d].view(1, Tk, H, E)
            v = cache[:Tk, d:].view(1, Tk, H, E)
            att = new(T, d)
            ops.attn_fwd(q, k, v, att.view(1, T, d), None, 1.0 / (E ** 0.5))
            hx = new(T, d)
            ops.linear_fwd(att, self.wT[nm + "attn.c_proj.weight"], hx, bias=m._wv(Wf, nm + "attn.c_proj.bias"),
                           residual=h, ld_res=d)
            c = new(T, d)
            ops.ln_fwd(hx, m._wv(Wf, nm + "ln_2.weight"), m._wv(Wf, nm + "ln_2.bias"), c)
            g = new(T, f)
            ops.linear_fwd(c, self.wT[nm + "mlp.c_fc.weight"], g, bias=m._wv(Wf, nm + "mlp.c_fc.bias"),
                           act=ops.ACT_GELU_NEW)
            h = new(T, d)
            ops.linear_fwd(g, self.wT[nm + "mlp.c_proj.weight"], h, bias=m._wv(Wf, nm + "mlp.c_proj.bias"),
                           residual=hx, ld_res=d)
        return h

    def _logits_into(self, hid_rows, out_rows, norm=None):
        """norm = (gamma, beta): hid_rows are pre-LayerNorm sums of the last layer (Performer, post-LN)"""
        m = self.m
        fs = getattr(self, "_fused", None)
        if fs is not None and out_rows is self.logits:
            # the step being captured ends in the sampler: projection + draw in ONE launch (emo_logits_sample)
            ops.logits_sample(hid_rows, m._wv(m.weights(), "dec_out_proj.weight"), m._wv(m._flat, "dec_out_proj.bias"),
                              self.logits, m.n_token, fs["t"], fs["p"], self.u_in, self._sampled[:self.B],
                              self._sampled[self.B:].view(torch.int32)[:self.B], greedy=fs["greedy"], banned=fs["banned"],
                              ln=norm)
            return
        ln = None
        if norm is not None:
            ln = (norm[0], norm[1], torch.empty(hid_rows.shape[0], m.d_model, dtype=self.dt, device=self.dev))
        ops.linear_fwd(hid_rows, m._wv(m.weights(), "dec_out_proj.weight"), out_rows[:, :m.n_token],
                       bias=m._wv(m._flat, "dec_out_proj.bias"), ln=ln)

    # ---- append a block of tokens to ONE sequence (primer / lead-sheet bar) -------------------------
    @torch.no_grad()
    def append(self, b, tokens, segs):
        """tokens / segs: python lists (or 1-D int64 tensors).  Returns the fp32 logits [V] after the last one."""
        self._sync_weights()
        m = self.m
        tok = torch.as_tensor(tokens, dtype=torch.int64).view(1, -1).to(self.dev)
        seg = torch.as_tensor(segs, dtype=torch.int64).view(1, -1).to(self.dev) if m.use_segment_emb else None
        T = tok.shape[1]
        pos0 = self.pos_host[b]
        if pos0 + T > self.max_len:
            raise RuntimeError("decode state is only valid up to max_len=%d positions" % self.max_len)
        d, H = m.d_model, m.n_head
        h = torch.empty(T, d, dtype=self.dt, device=self.dev)
        ops.embed_fwd(tok, seg, m._wv(m._flat, "token_emb.emb_lookup.weight"),
                      m._wv(m._flat, "segemb.emb_lookup.weight") if seg is not None else None,
                      m.pe.pe[pos0:] if m.use_pe else None, h, d ** 0.5)
        if self.is_performer:
            def favor(l, qkv, att):
                q3 = qkv.view(1, T, 3 * d)
                q, k, v = (q3[:, :, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
                st = self.state[l, b:b + 1]
                ops.favor_fwd(q, k, v, self.omegas[l], att.view(1, T, d), None, state_out=st, state_in=st)
            hid, norm = self._performer_rows(h, 1, T, favor)
        else:
            hid, norm = self._gpt2_rows(h, b, T, pos0), None
        self._logits_into(hid[T - 1:T], self.logits[b:b + 1], norm)
        self.pos_host[b] = pos0 + T
        self.pos[b] = pos0 + T
        return self.logits[b, :m.n_token]

    # ---- one token for every sequence ------------------------------------------------------------
    def _performer_step_body(self):
        m = self.m
        B, d, H = self.B, m.d_model, m.n_head
        if self.one_kernel:
            sl = m._sl
            ops.performer_decode_step(m.weights(), m._flat, self._offs, sl["token_emb.emb_lookup.weight"][0],
                                      sl["segemb.emb_lookup.weight"][0] if m.use_segment_emb else -1,
                                      sl["dec_out_proj.weight"][0], sl["dec_out_proj.bias"][0],
                                      m.pe.pe if m.use_pe else None, self.omegas, self.state, self.tok_in, self.seg_in,
                                      self.pos, self._scratch, self.logits, m.n_layer, B, m.n_token, d ** 0.5)
            if not m.use_pe:
                self.pos.add_(1)
            return
        h = torch.empty(B, d, dtype=self.dt, device=self.dev)
        ops.embed_rows(self.tok_in, self.seg_in if m.use_segment_emb else None, self.pos if m.use_pe else None,
                       m._wv(m._flat, "token_emb.emb_lookup.weight"),
                       m._wv(m._flat, "segemb.emb_lookup.weight") if m.use_segment_emb else None,
                       m.pe.pe if m.use_pe else None, h, d ** 0.5,
                       advance_pos=self.pos if m.use_pe else None)          # pos += 1 inside the kernel

        def favor(l, qkv, att):
            q, k, v = (qkv[:, i * d:(i + 1) * d].unflatten(-1, (H, E)) for i in range(3))
            ops.favor_step(q, k, v, self.omegas[l], self.state[l], att)
        hid, norm = self._performer_rows(h, B, 1, favor)
        self._logits_into(hid, self.logits, norm)
        if not m.use_pe:
            self.pos.add_(1)

    def _gpt2_step_body(self):
        """one new token for EVERY sequence of the ragged batch in static-shaped launches (graph-capturable): the
        projections see B rows, the attention reads each sequence's own cache length from the device"""
        m = self.m
        B, d, f, H = self.B, m.d_model, m.d_ff, m.n_head
        Wf = m._flat
        new = lambda *shape, dtype=self.dt: torch.empty(*shape, dtype=dtype, device=self.dev)
        self.pos_tok.copy_(self.pos)
        h = new(B, d)
        ops.embed_rows(self.tok_in, self.seg_in if m.use_segment_emb else None, self.pos if m.use_pe else None,
                       m._wv(Wf, "token_emb.emb_lookup.weight"),
                       m._wv(Wf, "segemb.emb_lookup.weight") if m.use_segment_emb else None,
                       m.pe.pe if m.use_pe else None, h, d ** 0.5, advance_pos=self.pos if m.use_pe else None)
        if not m.use_pe:
            self.pos.add_(1)
        scale = 1.0 / (E ** 0.5)
        for l in range(m.n_layer):
            nm = "transformer_decoder.%d." % l
            # ln_1 / ln_2 ride as the prologue of the projection that consumes them (in-kernel for these <= 8 rows):
            # 5 launches per layer
            a = new(B, d)
            qkv = new(B, 3 * d)
            ops.linear_fwd(h, self.wT[nm + "attn.c_attn.weight"], qkv, bias=m._wv(Wf, nm + "attn.c_attn.bias"),
                           ln=(m._wv(Wf, nm + "ln_1.weight"), m._wv(Wf, nm + "ln_1.bias"), a))
            att = new(B, d)
            ops.attn_decode_step(qkv, self.kv[l], self.pos_tok, att, scale)
            hx = new(B, d)
            ops.linear_fwd(att, self.wT[nm + "attn.c_proj.weight"], hx, bias=m._wv(Wf, nm + "attn.c_proj.bias"),
                           residual=h, ld_res=d)
            c = new(B, d)
            g = new(B, f)
            ops.linear_fwd(hx, self.wT[nm + "mlp.c_fc.weight"], g, bias=m._wv(Wf, nm + "mlp.c_fc.bias"), act=ops.ACT_GELU_NEW,
                           ln=(m._wv(Wf, nm + "ln_2.weight"), m._wv(Wf, nm + "ln_2.bias"), c))
            h = new(B, d)
            ops.linear_fwd(g, self.wT[nm + "mlp.c_proj.weight"], h, bias=m._wv(Wf, nm + "mlp.c_proj.bias"),
                           residual=hx, ld_res=d)
        self._logits_into(h, self.logits, None)

    def _step_body(self):
        if self.is_performer:
            self._performer_step_body()
        else:
            self._gpt2_step_body()

    @torch.no_grad()
    def step(self, tokens, segs):
        """tokens / segs: python lists of length B (one new token per sequence). Returns logits [B, V] (fp32,
        a view of a static buffer: consume before the next call)."""
        self._sync_weights()
        m = self.m
        if max(self.pos_host) + 1 > self.max_len:
            raise RuntimeError("decode state is only valid up to max_len=%d positions" % self.max_len)
        self._stage_inputs(tokens, segs, None)
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._step_body()
        for b in range(self.B):
            self.pos_host[b] += 1
        return self.logits[:, :m.n_token]

    def _stage_inputs(self, tokens, segs, us):
        """ONE small stream-ordered H2D per step from a pinned ring slot (rewritten only after its copy completed)"""
        i = self._ring_i
        self._ring_i = (i + 1) % len(self._ring)
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()
        host = self._ring[i]
        for b in range(self.B):
            host[0, b] = int(tokens[b])
            host[1, b] = int(segs[b])
        if us is not None:
            hu = host[2].view(torch.float32)
            for b in range(self.B):
                hu[b] = float(us[b])
        self._dev_in.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[i] = ev

    @torch.no_grad()
    def step_sample(self, tokens, segs, us, temperature, top_p, greedy=False, banned=None):
        """step() fused with the device sampler in ONE CUDA graph: tokens / segs / uniforms in by one H2D copy, the
        sampled ids (and the sampler's status words) back by one D2H copy.  Returns (ids, status) python lists;
        self.logits still holds the step's logits (a rejected draw is re-drawn from them by DeviceSampler)."""
        if not self.use_graph:
            raise RuntimeError("step_sample needs the graph path (use_graph=True)")
        self._sync_weights()
        if max(self.pos_host) + 1 > self.max_len:
            raise RuntimeError("decode state is only valid up to max_len=%d positions" % self.max_len)
        # banned: uint8 [B, V] device mask of inadmissible tokens (its ADDRESS is baked into the graph; the caller
        # updates the contents in place)
        # temperature: a float, or an fp32 [B] device tensor (one per sequence; its ADDRESS is baked into the graph)
        tkey = ("rows", temperature.data_ptr()) if torch.is_tensor(temperature) else float(temperature)
        self._temperature = temperature
        cfg = (tkey, float(top_p), bool(greedy), None if banned is None else banned.data_ptr())
        self._banned = banned
        if self.graph_sample is None or self.sample_cfg != cfg:
            self.sample_cfg = cfg
            self._capture(with_sampler=True)
        self._stage_inputs(tokens, segs, us if not greedy else [0.0] * self.B)
        self.graph_sample.replay()
        self._sampled_host.copy_(self._sampled, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        for b in range(self.B):
            self.pos_host[b] += 1
        ids = self._sampled_host[:self.B].tolist()
        st = self._sampled_host[self.B:].view(torch.int32)[:self.B].tolist()
        return ids, st

    def _capture(self, with_sampler=False):
        m = self.m
        m.weights()                                   # make sure the bf16 shadow exists before capture
        # the warm-up / capture-time launches advance the state: Performer prefix sums are restored below; the GPT-2
        # cache rows they write lie at or beyond `pos` and are rewritten when a real token reaches them
        state0, pos0 = (self.state.clone() if self.is_performer else None), self.pos.clone()
        if not self.is_performer and int(pos0.max()) + 3 > self.max_len:
            raise RuntimeError("capturing the GPT-2 step needs 3 free cache rows below max_len")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):                        # warm-up outside capture (lazy module / attribute init)
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        from . import _lib
        _lib.lib().emo_set_pdl(1 if self.use_pdl else 0)    # programmatic dependent launches inside the step graph
        try:
            with torch.cuda.graph(g):
                fused = (with_sampler and FUSED_SAMPLER and not self.one_kernel and self.dt == torch.bfloat16
                         and m.d_model == 512)
                if fused:
                    _, p, greedy, _ = self.sample_cfg
                    self._fused = {"t": self._temperature, "p": p, "greedy": greedy, "banned": self._banned}
                self._step_body()
                if with_sampler and not fused:
                    _, p, greedy, _ = self.sample_cfg
                    t = self._temperature
                    ops.sample(self.logits, m.n_token, t, p, self.u_in,
                               self._sampled[:self.B], self._sampled[self.B:].view(torch.int32)[:self.B], greedy=greedy,
                               banned=self._banned)
        finally:
            self._fused = None
            _lib.lib().emo_set_pdl(0)
        if state0 is not None:
            self.state.copy_(state0)                  # undo the warm-up / capture-time state advance
        self.pos.copy_(pos0)
        if with_sampler:
            self.graph_sample = g
        else:
            self.graph = g


class Stage1Decoder:
    """Incremental decode of the stage-1 lead-sheet model (PlainTransformer, Transformer-XL relative positions).

    The reference loop (stage1_compose/inference_utils.py:95-104) feeds one token and the hidden-state memory
    (n_layer + 1 tensors of the last mem_len layer inputs) to the model, which re-derives LayerNorm + K | V of EVERY
    memory row in every layer for every generated token (optimus_txl_decoder.py:702-722, 336-367).  K | V of a position
    never change at inference, so they are cached here; a step is 5 launches per layer on ONE row per sequence (the two
    LayerNorms ride as prologues of the projections that consume them), the attention kernel reads the last mem_len + 1
    cache rows and the per-layer table r[distance] = r_net(pos_emb(distance)), and the whole step (+ the sampler) is one
    CUDA graph.  Same arithmetic as PlainTransformer.generate with its memory (tests/test_decode_gpu.py)."""

    def __init__(self, model, batch=1, max_len=2560, use_graph=True, use_pdl=True):
        if model.training:
            raise RuntimeError("decode needs model.eval() (dropout off)")
        self.m, self.B, self.max_len = model, batch, max_len
        self.dev, self.dt = model._flat.device, model.compute_dtype
        self.mem_len = int(model.dec_mem_len)
        L, d = model.dec_n_layer, model.dec_d_model
        self.kv = torch.zeros(L, batch, max_len, 2 * d, dtype=self.dt, device=self.dev)
        self.pos = torch.zeros(batch, dtype=torch.int64, device=self.dev)
        self.pos_host = [0] * batch
        self.logits = torch.zeros(batch, model.ldv, dtype=torch.float32, device=self.dev)
        self.use_graph, self.use_pdl = bool(use_graph), bool(use_pdl)
        self.graph = None
        self._sample_graphs = {}          # (temperature, top_p, greedy) -> captured step + sampler graph
        self._ring = [torch.zeros(2, batch, dtype=torch.int64).pin_memory() for _ in range(8)]
        self._ring_ev = [None] * 8
        self._ring_i = 0
        self._dev_in = torch.zeros(2, batch, dtype=torch.int64, device=self.dev)      # tokens | uniforms (as fp32)
        self.tok_in = self._dev_in[0]
        self.u_in = self._dev_in[1].view(torch.float32)[:batch]
        self._sampled = torch.zeros(2 * batch, dtype=torch.int64, device=self.dev)
        self._sampled_host = torch.zeros(2 * batch, dtype=torch.int64).pin_memory()
        self.rtab = None

    def _sync_weights(self):
        m = self.m
        key = (m._weights_version(), m._flat.data_ptr())
        if key == getattr(self, "_wkey", None):
            return
        self._wkey = key
        self.graph = None
        self._sample_graphs = {}
        # r[distance] per layer: r_net applied to the sinusoid table; _pos_table(n) lists distances n-1 .. 0
        Wc = m.weights()
        M = self.mem_len
        pos = m._pos_table(M + 1, self.dev).flip(0).contiguous().to(self.dt)          # row = distance
        L, d = m.dec_n_layer, m.dec_d_model
        self.rtab = torch.empty(L, M + 1, d, dtype=self.dt, device=self.dev)
        for l in range(L):
            ops.linear_fwd(pos, m._wv(Wc, "decoder.layers.%d.dec_attn.r_net.weight" % l), self.rtab[l])

    def reset(self):
        self.pos.zero_()
        self.pos_host = [0] * self.B

    def _step_body(self):
        m = self.m
        B, d, f, L = self.B, m.dec_d_model, m.dec_d_ff, m.dec_n_layer
        Wc, Wf = m.weights(), m._flat
        new = lambda *shape, dtype=self.dt: torch.empty(*shape, dtype=dtype, device=self.dev)
        h = new(B, d)
        ops.embed_rows(self.tok_in, None, None, m._wv(Wf, "word_emb.emb_lookup.weight"), None, None, h, d ** 0.5)
        rw, rr = m._wv(Wf, "decoder.r_w_bias"), m._wv(Wf, "decoder.r_r_bias")
        scale = 1.0 / (E ** 0.5)
        for l in range(L):
            nm = "decoder.layers.%d." % l
            a, heads = new(B, d), new(B, 3 * d)
            ops.linear_fwd(h, m._wv(Wc, nm + "dec_attn.qkv_net.weight"), heads,
                           ln=(m._wv(Wf, nm + "dec_attn.layer_norm.weight"), m._wv(Wf, nm + "dec_attn.layer_norm.bias"), a))
            att = new(B, d)
            ops.relattn_decode_step(heads, self.kv[l], self.pos, self.rtab[l], rw, rr, self.mem_len, att, scale)
            h1 = new(B, d)
            ops.linear_fwd(att, m._wv(Wc, nm + "dec_attn.o_net.weight"), h1, residual=h, ld_res=d)
            c, ff = new(B, d), new(B, f)
            ops.linear_fwd(h1, m._wv(Wc, nm + "pos_ff.CoreNet.0.weight"), ff, bias=m._wv(Wf, nm + "pos_ff.CoreNet.0.bias"),
                           act=ops.ACT_RELU,
                           ln=(m._wv(Wf, nm + "pos_ff.layer_norm.weight"), m._wv(Wf, nm + "pos_ff.layer_norm.bias"), c))
            h = new(B, d)
            ops.linear_fwd(ff, m._wv(Wc, nm + "pos_ff.CoreNet.3.weight"), h, bias=m._wv(Wf, nm + "pos_ff.CoreNet.3.bias"),
                           residual=h1, ld_res=d)
        fs = getattr(self, "_fused", None)
        if fs is not None:         # the step being captured ends in the sampler: projection + draw in ONE launch
            ops.logits_sample(h, m._wv(Wc, "dec_out_proj.weight"), m._wv(Wf, "dec_out_proj.bias"), self.logits, m.vocab_size,
                              fs["t"], fs["p"], self.u_in, self._sampled[:self.B],
                              self._sampled[self.B:].view(torch.int32)[:self.B], greedy=fs["greedy"])
        else:
            ops.linear_fwd(h, m._wv(Wc, "dec_out_proj.weight"), self.logits[:, :m.vocab_size], bias=m._wv(Wf, "dec_out_proj.bias"))
        self.pos.add_(1)

    def _stage_inputs(self, tokens, us):
        i = self._ring_i
        self._ring_i = (i + 1) % len(self._ring)
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()
        host = self._ring[i]
        for b in range(self.B):
            host[0, b] = int(tokens[b])
        if us is not None:
            hu = host[1].view(torch.float32)
            for b in range(self.B):
                hu[b] = float(us[b])
        self._dev_in.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[i] = ev

    def _check_room(self):
        if max(self.pos_host) + 1 > self.max_len:
            raise RuntimeError("decode cache holds max_len=%d positions" % self.max_len)

    @torch.no_grad()
    def step(self, tokens):
        """tokens: list of B ids (one new token per sequence) -> fp32 logits [B, V] (a view of a static buffer)"""
        self._sync_weights()
        self._check_room()
        self._stage_inputs(tokens, None)
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._step_body()
        for b in range(self.B):
            self.pos_host[b] += 1
        return self.logits[:, :self.m.vocab_size]

    @torch.no_grad()
    def step_sample(self, tokens, us, temperature, top_p, greedy=False):
        """step() + the device sampler in one CUDA graph -> (ids, status) python lists; self.logits keeps the logits"""
        if not self.use_graph:
            raise RuntimeError("step_sample needs the graph path (use_graph=True)")
        self._sync_weights()
        self._check_room()
        cfg = (float(temperature), float(top_p), bool(greedy))
        if cfg not in self._sample_graphs:            # the loop alternates between two settings (first key draw / the rest)
            self._capture(cfg)
        self._stage_inputs(tokens, us if not greedy else [0.0] * self.B)
        self._sample_graphs[cfg].replay()
        self._sampled_host.copy_(self._sampled, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        for b in range(self.B):
            self.pos_host[b] += 1
        ids = self._sampled_host[:self.B].tolist()
        st = self._sampled_host[self.B:].view(torch.int32)[:self.B].tolist()
        return ids, st

    def _capture(self, sample_cfg=None):
        m = self.m
        m.weights()
        pos0 = self.pos.clone()
        if int(pos0.max()) + 3 > self.max_len:
            raise RuntimeError("capturing the step needs 3 free cache rows below max_len")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):                        # warm-up outside capture; the cache rows it writes lie at / beyond pos
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        from . import _lib
        _lib.lib().emo_set_pdl(1 if self.use_pdl else 0)
        try:
            with torch.cuda.graph(g):
                fused = sample_cfg is not None and FUSED_SAMPLER and self.dt == torch.bfloat16 and m.dec_d_model == 512
                if fused:
                    t, p, greedy = sample_cfg
                    self._fused = {"t": t, "p": p, "greedy": greedy}
                self._step_body()
                if sample_cfg is not None and not fused:
                    t, p, greedy = sample_cfg
                    ops.sample(self.logits, m.vocab_size, t, p, self.u_in, self._sampled[:self.B],
                               self._sampled[self.B:].view(torch.int32)[:self.B], greedy=greedy)
        finally:
            self._fused = None
            _lib.lib().emo_set_pdl(0)
        self.pos.copy_(pos0)
        if sample_cfg is not None:
            self._sample_graphs[sample_cfg] = g
        else:
            self.graph = g
