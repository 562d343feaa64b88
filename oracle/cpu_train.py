"""CPU baseline of the stage-2 Performer TRAIN STEP (reference stage2_accompaniment/train.py:58-81:
forward, CE, backward, clip_grad_norm_(0.5), Adam) restated over the oracle.  TEST INFRASTRUCTURE:
imported only by bench.py's `cpu_baseline` / `--impl reference` legs and by tests.

The reference's own module cannot run on the GPU box (fast_transformers is not installable, and
/root/reference does not travel), so the timed thing is the oracle port: pure torch fp32 on the
host cores, dropout 0.1 as in training (emopia_finetune.yaml), Omega redrawn every forward like
fast-transformers' Favor does."""
import time
import torch
import torch.nn.functional as F

from . import performer_oracle as PO


def synthetic_batch(V, B, T, seed):
    """tokens uniform over [0, V-2], seg ~ Bernoulli(0.5), targets = inputs shifted by one
    with PAD (= V-1) wherever the position is not on the Full track (mimics dataloader.py:127-144)."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, V - 1, (B, T), generator=g)
    seg = torch.randint(0, 2, (B, T), generator=g)
    tgt = torch.where(seg == 1, torch.roll(tok, -1, 1), torch.full_like(tok, V - 1))
    return tok, seg, tgt


class CpuPerformerTrainer:
    def __init__(self, V=329, n_layer=12, n_head=8, d_model=512, d_ff=2048, dropout=0.1, lr=1e-4, seed=0):
        self.V, self.L, self.H, self.d, self.p = V, n_layer, n_head, d_model, dropout
        shapes = PO.performer_state_shapes(V, n_layer, d_model, d_ff)
        sd = PO.seeded_state(shapes, seed)
        self.params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        self.sd = dict(self.params)
        self.sd["pe.pe"] = PO.sinusoid_pe(12000, d_model)
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)

    def step(self, tok, seg, tgt):
        omegas = [PO.draw_omega(64, 64) for _ in range(self.L)]
        drop = (lambda t, site: F.dropout(t, self.p, True)) if self.p > 0 else None
        x = PO.embed(tok, seg, self.sd, self.d)
        if drop is not None:
            x = drop(x, "emb")
        for l in range(self.L):
            x = PO.encoder_layer(x, self.sd, "transformer_decoder.decoder_layers.%d" % l, omegas[l], self.H,
                                 dropout=drop)
        logits = F.linear(x, self.sd["dec_out_proj.weight"], self.sd["dec_out_proj.bias"])
        loss = PO.ce_loss(logits, tgt, self.V - 1)
        self.opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(self.params.values()), 0.5)
        self.opt.step()
        return float(loss)


def time_cpu_train(V=329, B=1, T=2048, steps=3, warmup=1, threads=None, n_layer=12):
    """Returns dict(value tokens/s, ms_per_step, cores, sample)."""
    import os
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = CpuPerformerTrainer(V=V, n_layer=n_layer)
    tok, seg, tgt = synthetic_batch(V, B, T, 0)
    for _ in range(warmup):
        tr.step(tok, seg, tgt)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(tok, seg, tgt)
        ts.append(time.perf_counter() - t0)
    tot = sum(ts)
    return {"value": B * T * steps / tot, "ms_per_step": 1e3 * tot / steps, "cores": cores,
            "sample": "%d step(s) of B=%d x T=%d (V=%d, %d layers, fp32 torch on host, dropout 0.1, Omega redrawn "
                      "per forward); %d warm-up" % (steps, B, T, V, n_layer, warmup)}


# ------------------------------------------------------------------------------------------------
# CPU twins of the other two train steps (BASELINE.json configs[2] and configs[0]); eval-mode math of the oracle
# (no dropout: dropout is a negligible share of a CPU step) + CE + backward + clip_grad_norm_(0.5) + Adam
# ------------------------------------------------------------------------------------------------
def _time_steps(step, units, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    tot = time.perf_counter() - t0
    return units * steps / tot, 1e3 * tot / steps


def time_cpu_train_gpt2(V=372, B=1, T=2048, steps=1, warmup=1, threads=None, n_layer=12):
    """stage2_accompaniment/train.py:58-81 with -m gpt2 (music_gpt2.py + HF GPT2Block) over oracle/gpt2_oracle.py"""
    import os
    from . import gpt2_oracle as GO
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = PO.seeded_state(GO.gpt2_state_shapes(V, n_layer), 0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    full = dict(params)
    full["pe.pe"] = PO.sinusoid_pe(12000, 512)
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    tok, seg, tgt = synthetic_batch(V, B, T, 0)

    def step():
        logits = GO.gpt2_forward(full, tok, seg, n_layer, 8, 512)
        loss = PO.ce_loss(logits, tgt, V - 1)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 0.5)
        opt.step()
    v, ms = _time_steps(step, B * T, steps, warmup)
    return {"value": v, "ms_per_step": ms, "cores": cores,
            "sample": "%d step(s) of B=%d x T=%d (REMI V=%d, %d GPT-2 blocks, fp32 torch on host, O(T^2) attention); %d warm-up"
                      % (steps, B, T, V, n_layer, warmup)}


def time_cpu_train_stage1(V=216, B=1, T=512, steps=2, warmup=1, threads=None, n_layer=12):
    """stage1_compose/train.py:48-65 (PlainTransformer, mem_len 0 in training) over oracle/txl_oracle.py, which is
    pinned against the unmodified reference model (oracle/validate_against_reference.py)"""
    import os
    from . import txl_oracle as TO
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = PO.seeded_state(TO.txl_state_shapes(V, n_layer), 0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(0, V - 1, (T, B), generator=g)
    tgt = torch.roll(tok, -1, 0)

    def step():
        logits, _ = TO.txl_forward(params, tok, None, n_layer, 8, 512, 0)
        loss = F.cross_entropy(logits.view(-1, V), tgt.view(-1), ignore_index=V - 1)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 0.5)
        opt.step()
    v, ms = _time_steps(step, B * T, steps, warmup)
    return {"value": v, "ms_per_step": ms, "cores": cores,
            "sample": "%d step(s) of T=%d x B=%d (functional lead-sheet V=%d, %d Transformer-XL layers, fp32 torch on host); "
                      "%d warm-up" % (steps, T, B, V, n_layer, warmup)}
