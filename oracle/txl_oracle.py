"""CPU oracle: stage-1 PlainTransformer (Transformer-XL style rel-pos decoder).  TEST INFRA.

Restates reference stage1_compose/model/plain_transformer.py:51-93,
optimus_txl_decoder.py:8-24 (PositionalEmbedding), :47-61 (PositionwiseFF, pre-LN),
:305-387 (RelPartialLearnableMultiHeadAttn), :702-748 (_update_mems), :750-925 (_forward,
attn_type 0, use_segment_emb=False).  Layout [T, B, .] like the reference.  Eval mode
(dropout off; the post-dropout renormalisation P/(sum P + 1e-8) at :362-363 is kept).
PINNED against the reference module by oracle/validate_against_reference.py.
"""
import torch
import torch.nn.functional as F

from .performer_oracle import LN_EPS


def rel_pos_table(klen, d_model, dtype=torch.float32):
    """pos_emb for distances klen-1 .. 0: [sin | cos] halves (optimus_txl_decoder.py:16-19)."""
    inv_freq = 1 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model))
    pos_seq = torch.arange(klen - 1, -1, -1.0, dtype=dtype)
    s = torch.outer(pos_seq, inv_freq.to(dtype))
    return torch.cat([s.sin(), s.cos()], dim=-1)  # [klen, d]


def rel_attention(h, mem, sd, prefix, r_w_bias, r_r_bias, n_head):
    """h [q,B,d], mem [m,B,d] or None -> h + o_net(attn) (pre-LN residual form)."""
    qlen, B, d = h.shape
    dh = d // n_head
    cat = h if mem is None or mem.numel() == 0 else torch.cat([mem, h], 0)
    klen = cat.shape[0]
    mlen = klen - qlen
    lnw, lnb = sd[prefix + ".dec_attn.layer_norm.weight"], sd[prefix + ".dec_attn.layer_norm.bias"]
    heads = F.linear(F.layer_norm(cat, (d,), lnw, lnb, LN_EPS), sd[prefix + ".dec_attn.qkv_net.weight"])
    q, k, v = heads.chunk(3, dim=-1)
    q = q[-qlen:].view(qlen, B, n_head, dh)
    k = k.view(klen, B, n_head, dh)
    v = v.view(klen, B, n_head, dh)
    r = F.linear(rel_pos_table(klen, d, h.dtype), sd[prefix + ".dec_attn.r_net.weight"]).view(klen, n_head, dh)
    AC = torch.einsum("ibnd,jbnd->ijbn", q + r_w_bias, k)
    BDfull = torch.einsum("ibnd,pnd->ipbn", q + r_r_bias, r)       # p indexes distance klen-1-p
    i = torch.arange(qlen)[:, None]
    j = torch.arange(klen)[None, :]
    dist = i + mlen - j                                            # >= 0 where visible
    p = (klen - 1 - dist).clamp(0, klen - 1)
    BD = torch.gather(BDfull, 1, p[:, :, None, None].expand(-1, -1, B, n_head))
    score = (AC + BD) * (1.0 / dh ** 0.5)
    score = score.masked_fill((dist < 0)[:, :, None, None], float("-inf"))
    prob = F.softmax(score, dim=1)
    prob = prob / (prob.sum(dim=1, keepdim=True) + 1e-8)
    vec = torch.einsum("ijbn,jbnd->ibnd", prob, v).reshape(qlen, B, d)
    return h + F.linear(vec, sd[prefix + ".dec_attn.o_net.weight"])


def pos_ff(h, sd, prefix):
    d = h.shape[-1]
    y = F.layer_norm(h, (d,), sd[prefix + ".pos_ff.layer_norm.weight"], sd[prefix + ".pos_ff.layer_norm.bias"], LN_EPS)
    y = F.relu(F.linear(y, sd[prefix + ".pos_ff.CoreNet.0.weight"], sd[prefix + ".pos_ff.CoreNet.0.bias"]))
    y = F.linear(y, sd[prefix + ".pos_ff.CoreNet.3.weight"], sd[prefix + ".pos_ff.CoreNet.3.bias"])
    return h + y


def txl_forward(sd, tokens, mems, n_layer, n_head, d_model, mem_len, taps=None):
    """PlainTransformer.forward: tokens [T,B] int64, mems None/() or list of n_layer+1 [m,B,d].
    Returns (logits [T,B,V], new_mems or None)."""
    h = sd["word_emb.emb_lookup.weight"][tokens] * (d_model ** 0.5)
    qlen = tokens.shape[0]
    have_mems = mem_len > 0
    if have_mems and (mems is None or len(mems) == 0):
        mems = [torch.empty(0, tokens.shape[1], d_model, dtype=h.dtype) for _ in range(n_layer + 1)]
    hids = [h]
    for l in range(n_layer):
        mem = mems[l] if have_mems else None
        p = "decoder.layers.%d" % l
        h = rel_attention(h, mem, sd, p, sd["decoder.r_w_bias"], sd["decoder.r_r_bias"], n_head)
        h = pos_ff(h, sd, p)
        hids.append(h)
        if taps is not None:
            taps.append(h)
    logits = F.linear(h, sd["dec_out_proj.weight"], sd["dec_out_proj.bias"])
    new_mems = None
    if have_mems:
        mlen = mems[0].shape[0]
        end = mlen + qlen
        beg = max(0, end - mem_len)
        new_mems = [torch.cat([mems[i], hids[i]], 0)[beg:end] for i in range(n_layer + 1)]
    return logits, new_mems


def txl_generate(sd, tokens, mems, n_layer, n_head, d_model, mem_len):
    """PlainTransformer.generate: logits of the last row for batch element 0."""
    logits, new_mems = txl_forward(sd, tokens, mems, n_layer, n_head, d_model, mem_len)
    return logits[-1, 0, :], new_mems


def txl_state_shapes(vocab, n_layer, d_model=512, d_ff=2048, n_head=8):
    dh = d_model // n_head
    shapes = {
        "word_emb.emb_lookup.weight": (vocab, d_model),
        "decoder.r_w_bias": (n_head, dh),
        "decoder.r_r_bias": (n_head, dh),
        "dec_out_proj.weight": (vocab, d_model),
        "dec_out_proj.bias": (vocab,),
    }
    for l in range(n_layer):
        p = "decoder.layers.%d." % l
        shapes[p + "dec_attn.qkv_net.weight"] = (3 * d_model, d_model)
        shapes[p + "dec_attn.r_net.weight"] = (d_model, d_model)
        shapes[p + "dec_attn.o_net.weight"] = (d_model, d_model)
        shapes[p + "dec_attn.layer_norm.weight"] = (d_model,)
        shapes[p + "dec_attn.layer_norm.bias"] = (d_model,)
        shapes[p + "pos_ff.CoreNet.0.weight"] = (d_ff, d_model)
        shapes[p + "pos_ff.CoreNet.0.bias"] = (d_ff,)
        shapes[p + "pos_ff.CoreNet.3.weight"] = (d_model, d_ff)
        shapes[p + "pos_ff.CoreNet.3.bias"] = (d_model,)
        shapes[p + "pos_ff.layer_norm.weight"] = (d_model,)
        shapes[p + "pos_ff.layer_norm.bias"] = (d_model,)
    return shapes
