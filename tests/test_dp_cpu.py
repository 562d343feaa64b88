"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic: dp.GradSync's collectives and
the loss / gradient semantics they implement -- the summed gradient of   local_loss_sum / GLOBAL_count   over
ranks must equal the gradient of the 1-GPU mean CE over the concatenated batch (reference
stage2_accompaniment/train.py:71-79 at N x batch).  The model math here is the oracle (checker only)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FlatStub:
    """what GradSync touches of a FlatModule: the flat fp32 parameter / gradient buffers"""

    def __init__(self, n):
        self._flat = torch.zeros(n)
        self._flat_grad = torch.zeros(n)
        self._lp_version = 0


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:          # surface the failure instead of letting the parent time out
        q.put((rank, "%s: %s" % (type(e).__name__, e), None))
        raise


def _worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    from emo_disentanger_b200 import dp
    from oracle import performer_oracle as PO
    from oracle.cpu_train import synthetic_batch
    r, local, w = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    V, L, B, T = 40, 1, 2, 24
    shapes = PO.performer_state_shapes(V, L)
    names = sorted(shapes)
    sd = PO.seeded_state(shapes, 5)
    sd["pe.pe"] = PO.sinusoid_pe(64, 512)
    om = [torch.randn(64, 64, generator=torch.Generator().manual_seed(9))]
    n = sum(sd[k].numel() for k in names)
    model = _FlatStub(n)
    sync = dp.GradSync(model)
    assert sync.world == world

    # broadcast_params: every rank ends with rank 0's parameters, bf16 shadow invalidated
    model._flat.fill_(float(rank + 1))
    sync.broadcast_params()
    assert float(model._flat.min()) == 1.0 and float(model._flat.max()) == 1.0 and model._lp_version == -1

    def grads(tok, seg, tgt, count):
        p = {k: sd[k].clone().requires_grad_(True) for k in names}
        full = dict(p); full["pe.pe"] = sd["pe.pe"]
        logits = PO.performer_forward(full, tok, seg, om, L, 8, 512)
        lsum = torch.nn.functional.cross_entropy(logits.reshape(-1, V), tgt.reshape(-1), ignore_index=V - 1, reduction="sum")
        (lsum / count).backward()
        return torch.cat([p[k].grad.reshape(-1) for k in names]), float(lsum)

    # each rank: its own shard (rank-strided samples of the global batch)
    tok, seg, tgt = synthetic_batch(V, B * world, T, 77)
    mine = slice(rank, None, world)
    local_count = (tgt[mine] != V - 1).sum().float().view(1)
    count = sync.count_allreduce(local_count.clone())
    assert float(count) == float((tgt != V - 1).sum())
    g, lsum = grads(tok[mine], seg[mine], tgt[mine], count)
    model._flat_grad.copy_(g)
    sync.allreduce_grads()
    acc = sync.allreduce_stats(torch.tensor([float(count), lsum, 0.0]))
    # single-process result over the WHOLE batch
    g_ref, lsum_ref = grads(tok, seg, tgt, (tgt != V - 1).sum().float())
    err = float((model._flat_grad - g_ref).abs().max() / g_ref.abs().max())
    # bucketed path: "layers" reduced early from the backward's call-backs (reverse order, async), the rest at the end
    reduced = model._flat_grad.clone()
    cuts = [n // 7, n // 3, n // 2, (3 * n) // 4]           # three layer buckets in the middle; head and tail are "the rest"
    model.layer_grad_range = lambda l: (cuts[l], cuts[l + 1])
    model._flat_grad.copy_(g)
    sync.begin_step(last_micro_batch=False)                   # accumulating micro-batch: the call-backs must not reduce
    for l in (2, 1, 0):
        model.grad_hook(l)
    assert torch.equal(model._flat_grad, g) and not sync._work
    sync.begin_step(last_micro_batch=True)
    for l in (2, 1, 0):
        model.grad_hook(l)
    assert len(sync._work) == 3
    sync.allreduce_grads()
    assert not sync._work and not sync._ranges
    err = max(err, float((model._flat_grad - reduced).abs().max()))     # same sums, bucket by bucket
    q.put((rank, err, abs(float(acc[1]) - lsum_ref) / lsum_ref))
    dist.destroy_process_group()


def test_gradsync_world2_gloo_matches_single_process_global_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gerr, lerr in res:
        assert not isinstance(gerr, str), gerr
        assert gerr < 1e-5, (rank, gerr)
        assert lerr < 1e-6, (rank, lerr)


def test_world1_is_no_communication():
    sys.path.insert(0, ROOT)
    from emo_disentanger_b200 import dp
    m = _FlatStub(8)
    s = dp.GradSync(m)
    assert s.world == 1
    m._flat_grad.fill_(2.0)
    s.allreduce_grads()
    assert float(m._flat_grad.sum()) == 16.0
    c = torch.tensor([3.0])
    assert s.count_allreduce(c) is c
