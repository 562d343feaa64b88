"""CPU oracle: stage-2 MusicGPT2.  TEST INFRASTRUCTURE.

Restates reference stage2_accompaniment/model/music_gpt2.py:70-103 (embedding front-end shared
with the Performer, 12x HF GPT2Block, no ln_f, dec_out_proj) and the third-party HF
``GPT2Block`` (README.md:19 pins transformers==4.28.0): pre-LN,
Conv1D y = x W + b with W [in,out], eager causal softmax attention scaled by 1/sqrt(dh),
gelu_new MLP.  Pinned against HF 5.5.0 GPT2Block by oracle/validate_against_reference.py.
"""
import math
import torch
import torch.nn.functional as F

from .performer_oracle import embed, LN_EPS


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def conv1d(x, w, b):
    return x @ w + b


def causal_softmax_attention(q, k, v):
    """q,k,v [N,H,L,dh] -> [N,H,L,dh]; mask value = finfo.min as HF eager attention."""
    L = q.shape[2]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    mask = torch.tril(torch.ones(L, L, dtype=torch.bool))
    s = torch.where(mask, s, torch.full([], torch.finfo(s.dtype).min, dtype=s.dtype))
    return F.softmax(s, dim=-1) @ v


def gpt2_block(x, sd, prefix, n_head, taps=None):
    N, L, d = x.shape
    h = F.layer_norm(x, (d,), sd[prefix + ".ln_1.weight"], sd[prefix + ".ln_1.bias"], LN_EPS)
    qkv = conv1d(h, sd[prefix + ".attn.c_attn.weight"], sd[prefix + ".attn.c_attn.bias"])
    q, k, v = qkv.split(d, dim=2)
    sh = lambda t: t.view(N, L, n_head, d // n_head).permute(0, 2, 1, 3)
    a = causal_softmax_attention(sh(q), sh(k), sh(v)).permute(0, 2, 1, 3).reshape(N, L, d)
    x = x + conv1d(a, sd[prefix + ".attn.c_proj.weight"], sd[prefix + ".attn.c_proj.bias"])
    h = F.layer_norm(x, (d,), sd[prefix + ".ln_2.weight"], sd[prefix + ".ln_2.bias"], LN_EPS)
    h = gelu_new(conv1d(h, sd[prefix + ".mlp.c_fc.weight"], sd[prefix + ".mlp.c_fc.bias"]))
    x = x + conv1d(h, sd[prefix + ".mlp.c_proj.weight"], sd[prefix + ".mlp.c_proj.bias"])
    if taps is not None:
        taps.append(x)
    return x


def gpt2_forward(sd, tokens, seg, n_layer, n_head, d_model, keep_last_only=False, taps=None):
    x = embed(tokens, seg, sd, d_model)
    if taps is not None:
        taps.append(x)
    for l in range(n_layer):
        x = gpt2_block(x, sd, "transformer_decoder.%d" % l, n_head, taps=taps)
    logits = F.linear(x, sd["dec_out_proj.weight"], sd["dec_out_proj.bias"])
    if keep_last_only:
        logits = logits[:, -1, :]
    return logits


def gpt2_state_shapes(n_token, n_layer, d_model=512, d_ff=2048, n_seg=2):
    shapes = {
        "token_emb.emb_lookup.weight": (n_token, d_model),
        "segemb.emb_lookup.weight": (n_seg, d_model),
        "dec_out_proj.weight": (n_token, d_model),
        "dec_out_proj.bias": (n_token,),
    }
    for l in range(n_layer):
        p = "transformer_decoder.%d." % l
        for nm in ("ln_1", "ln_2"):
            shapes[p + nm + ".weight"] = (d_model,)
            shapes[p + nm + ".bias"] = (d_model,)
        shapes[p + "attn.c_attn.weight"] = (d_model, 3 * d_model)
        shapes[p + "attn.c_attn.bias"] = (3 * d_model,)
        shapes[p + "attn.c_proj.weight"] = (d_model, d_model)
        shapes[p + "attn.c_proj.bias"] = (d_model,)
        shapes[p + "mlp.c_fc.weight"] = (d_model, d_ff)
        shapes[p + "mlp.c_fc.bias"] = (d_ff,)
        shapes[p + "mlp.c_proj.weight"] = (d_ff, d_model)
        shapes[p + "mlp.c_proj.bias"] = (d_model,)
    return shapes
