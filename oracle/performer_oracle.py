"""CPU oracle: stage-2 MusicPerformer (FAVOR+ causal linear attention).  TEST INFRASTRUCTURE.

Restates, as plain functions over a reference-keyed state dict:
  * model glue      -- reference stage2_accompaniment/model/music_performer.py:50-81,
                       transformer_helpers.py:43-87, fast_transformer_decoder.py:54-74
  * third-party math (fast-transformers 0.4.0, NOT under /root/reference, un-pinned:
    README.md:13-16) -- SURVEY.md Appendix A.1-A.5:
      Favor.forward                      feature_maps/fourier_features.py
      CausalLinearAttention.forward      attention/causal_linear_attention.py
      causal_dot_product (CPU kernel)    causal_product/causal_product_cpu.cpp
      AttentionLayer / TransformerEncoderLayer   attention/attention_layer.py, transformers.py
  parity unpinned for the third-party part (no upstream source / goldens available offline).
"""
import math
import torch
import torch.nn.functional as F

N_FEAT_DEFAULT = 128
EPS_ATTN = 1e-6
LN_EPS = 1e-5


def sinusoid_pe(max_pos, d_embed, dtype=torch.float32):
    # transformer_helpers.py:48-54 -- interleaved sin/cos, buffer shape [max_pos, 1, d]
    pe = torch.zeros(max_pos, d_embed)
    position = torch.arange(0, max_pos, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_embed, 2).float() * (-math.log(10000.0) / d_embed))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(1).to(dtype)


def draw_omega(n_rows, n_cols, generator=None, orthogonal=False, dtype=torch.float32):
    """Favor.new_feature_map: omega [query_dims, n_dims//2].

    fast-transformers RandomFourierFeatures.new_feature_map: ``omega.normal_()`` unless
    ``orthogonal=True`` (then block-wise QR scaled by row norms of the Gaussian, App. A.1).
    The reference passes no ``orthogonal`` kwarg (fast_transformer_decoder.py:28-31)."""
    if not orthogonal:
        return torch.randn(n_rows, n_cols, generator=generator, dtype=dtype)
    w = torch.empty(n_rows, n_cols, dtype=dtype)
    start = 0
    while start < n_cols:
        end = min(start + n_rows, n_cols)
        block = torch.randn(n_rows, n_rows, generator=generator, dtype=dtype)
        norms = torch.sqrt(torch.einsum("ab,ab->a", block, block))
        Q, _ = torch.linalg.qr(block)
        w[:, start:end] = Q[:, : end - start] * norms[None, : end - start]
        start += n_rows
    return w


def favor_features(x, omega, n_dims=N_FEAT_DEFAULT):
    """Favor.forward (stabilize=False): x [..., E] -> phi [..., n_dims]."""
    E = x.shape[-1]
    softmax_temp = 1.0 / math.sqrt(E)
    x = x * math.sqrt(softmax_temp)
    norm_sq = (x * x).sum(-1, keepdim=True)
    u = x @ omega
    offset = norm_sq * 0.5 + 0.5 * math.log(n_dims)
    return torch.cat([torch.exp(u - offset), torch.exp(-u - offset)], dim=-1)


def causal_dot_product_seq(Q, K, V):
    """Sequential statement of causal_product_cpu.cpp: Q,K [N,H,L,M], V [N,H,L,D]."""
    N, H, L, M = Q.shape
    D = V.shape[-1]
    out = torch.zeros(N, H, L, D, dtype=Q.dtype)
    kv = torch.zeros(N, H, M, D, dtype=Q.dtype)
    for l in range(L):
        kv = kv + K[:, :, l, :, None] * V[:, :, l, None, :]
        out[:, :, l] = torch.einsum("nhm,nhmd->nhd", Q[:, :, l], kv)
    return out


def causal_dot_product_chunked(Q, K, V, chunk=128):
    """Same result (up to summation order), O(L) memory.  Q,K [N,H,L,M], V [N,H,L,D]."""
    N, H, L, M = Q.shape
    D = V.shape[-1]
    outs = []
    state = torch.zeros(N, H, M, D, dtype=Q.dtype)
    for s in range(0, L, chunk):
        q, k, v = Q[:, :, s:s + chunk], K[:, :, s:s + chunk], V[:, :, s:s + chunk]
        c = q.shape[2]
        a = torch.einsum("nhlm,nhjm->nhlj", q, k)
        a = a * torch.tril(torch.ones(c, c, dtype=Q.dtype))
        o = torch.einsum("nhlj,nhjd->nhld", a, v) + torch.einsum("nhlm,nhmd->nhld", q, state)
        outs.append(o)
        state = state + torch.einsum("nhjm,nhjd->nhmd", k, v)
    return torch.cat(outs, dim=2)


def causal_linear_attention(q, k, v, omega, n_dims=N_FEAT_DEFAULT, sequential=False):
    """CausalLinearAttention.forward, q,k,v [N,L,H,E] -> [N,L,H,D]; returns (out, denom)."""
    Q = favor_features(q, omega, n_dims)
    K = favor_features(k, omega, n_dims)
    den = torch.einsum("nlhi,nlhi->nlh", Q, K.cumsum(1)) + EPS_ATTN
    Qp, Kp, Vp = (t.permute(0, 2, 1, 3).contiguous() for t in (Q, K, v))
    fn = causal_dot_product_seq if sequential else causal_dot_product_chunked
    out = fn(Qp, Kp, Vp).permute(0, 2, 1, 3)
    return out * (1.0 / den)[..., None], den


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def encoder_layer(x, sd, prefix, omega, n_head, n_dims=N_FEAT_DEFAULT, activation="relu",
                  dropout=None, taps=None):
    """fast_transformers.transformers.TransformerEncoderLayer.forward (post-LN), eval or with an
    injected ``dropout(t, site)`` callable."""
    drop = dropout if dropout is not None else (lambda t, site: t)
    N, L, d = x.shape
    a = prefix + ".attention"
    q = _lin(x, sd, a + ".query_projection").view(N, L, n_head, -1)
    k = _lin(x, sd, a + ".key_projection").view(N, L, n_head, -1)
    v = _lin(x, sd, a + ".value_projection").view(N, L, n_head, -1)
    att, _ = causal_linear_attention(q, k, v, omega, n_dims)
    att = _lin(att.reshape(N, L, d), sd, a + ".out_projection")
    x = x + drop(att, "attn")
    x = F.layer_norm(x, (d,), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"], LN_EPS)
    act = F.relu if activation == "relu" else F.gelu
    y = drop(act(_lin(x, sd, prefix + ".linear1")), "ffn1")
    y = drop(_lin(y, sd, prefix + ".linear2"), "ffn2")
    out = F.layer_norm(x + y, (d,), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], LN_EPS)
    if taps is not None:
        taps.append(out)
    return out


def embed(tokens, seg, sd, d_model, use_pe=True):
    """TokenEmbedding x sqrt(d) (+ segemb x sqrt(d)) + pe  (music_performer.py:51-62)."""
    scale = d_model ** 0.5
    x = sd["token_emb.emb_lookup.weight"][tokens] * scale
    if seg is not None and "segemb.emb_lookup.weight" in sd:
        x = x + sd["segemb.emb_lookup.weight"][seg] * scale
    if use_pe:
        x = x + sd["pe.pe"][: tokens.shape[1]].permute(1, 0, 2)
    return x


def performer_forward(sd, tokens, seg, omegas, n_layer, n_head, d_model,
                      n_dims=N_FEAT_DEFAULT, keep_last_only=False, taps=None):
    """MusicPerformer.forward in eval mode; omegas: list of n_layer [E, n_dims//2] tensors."""
    x = embed(tokens, seg, sd, d_model)
    if taps is not None:
        taps.append(x)
    for l in range(n_layer):
        x = encoder_layer(x, sd, "transformer_decoder.decoder_layers.%d" % l, omegas[l], n_head,
                          n_dims, taps=taps)
    logits = F.linear(x, sd["dec_out_proj.weight"], sd["dec_out_proj.bias"])
    if keep_last_only:
        logits = logits[:, -1, :]
    return logits


def ce_loss(logits, tgt, pad_idx):
    """compute_loss (music_performer.py:72-81): mean CE over targets != n_token-1."""
    return F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), tgt.reshape(-1),
                           ignore_index=pad_idx, reduction="mean")


def performer_state_shapes(n_token, n_layer, d_model=512, d_ff=2048, n_seg=2, max_pos=12000):
    shapes = {
        "token_emb.emb_lookup.weight": (n_token, d_model),
        "segemb.emb_lookup.weight": (n_seg, d_model),
        "dec_out_proj.weight": (n_token, d_model),
        "dec_out_proj.bias": (n_token,),
    }
    for l in range(n_layer):
        p = "transformer_decoder.decoder_layers.%d." % l
        for nm in ("query", "key", "value", "out"):
            shapes[p + "attention.%s_projection.weight" % nm] = (d_model, d_model)
            shapes[p + "attention.%s_projection.bias" % nm] = (d_model,)
        shapes[p + "linear1.weight"] = (d_ff, d_model)
        shapes[p + "linear1.bias"] = (d_ff,)
        shapes[p + "linear2.weight"] = (d_model, d_ff)
        shapes[p + "linear2.bias"] = (d_model,)
        for nm in ("norm1", "norm2"):
            shapes[p + nm + ".weight"] = (d_model,)
            shapes[p + nm + ".bias"] = (d_model,)
    return shapes


def seeded_state(shapes, seed, std=0.02, ln_std=0.02, bias_std=0.02, extra=None):
    """Deterministic synthetic weights (CPU generator) shared by goldens and GPU tests.
    Larger than the reference's N(0,0.01) init on purpose: biases / LN offsets are non-zero so
    every term of the path is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("layer_norm.weight") \
                or k.endswith("ln_1.weight") or k.endswith("ln_2.weight"):
            sd[k] = 1.0 + ln_std * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = bias_std * torch.randn(shp, generator=g)
        else:
            sd[k] = std * torch.randn(shp, generator=g)
    if extra:
        sd.update(extra)
    return sd
