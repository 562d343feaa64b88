"""Host-side logic that needs no GPU: bf16-shadow invalidation, reference-ordered optimizer checkpoints, the
rank-sharded epoch order of the training scripts."""
import json
import os

import pytest
import torch

from emo_disentanger_b200.engine import FlatModule
from emo_disentanger_b200.optim import FusedAdam
from emo_disentanger_b200.stage1 import PlainTransformer
from emo_disentanger_b200.stage2 import MusicGPT2, MusicPerformer

HERE = os.path.dirname(os.path.abspath(__file__))
ORDER = json.load(open(os.path.join(HERE, "golden", "param_order.json")))
KW = dict(n_token=20, n_layer=2, n_head=8, d_model=512, d_ff=2048, d_embed=512, use_segment_emb=True, n_segment_types=2)


def _models():
    return {"performer": MusicPerformer(favor_feature_dims=128, **KW), "gpt2": MusicGPT2(**KW),
            "stage1": PlainTransformer(512, 20, 2, 8, 512, 2048, 0, 64, pad_index=19, pre_lnorm=True)}


class _Tiny(FlatModule):
    def __init__(self):
        super().__init__()
        self._add_param("a.w", (4, 4), lambda v: v.normal_())
        self._add_param("b", (3,), lambda v: v.zero_())
        self._finalize()


def test_weights_version_moves_with_external_optimizer_and_load_state_dict():
    """every Parameter is bound with `p.data = view`, so `_flat._version` stays put when
    torch.optim.Adam.step() or load_state_dict() rewrite the masters; the shadow-refresh key must still move."""
    m = _Tiny()
    v0 = m._weights_version()
    opt = torch.optim.Adam(m.parameters(), 1e-2)
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    v1 = m._weights_version()
    assert v1 > v0
    m.load_state_dict({k: v + 1 for k, v in m.state_dict().items()})
    assert m._weights_version() > v1
    m.mark_lp_fresh()
    assert m._lp_version == m._weights_version()


@pytest.mark.parametrize("which", ["performer", "gpt2", "stage1"])
def test_reference_parameter_order(which):
    m = _models()[which]
    assert m.reference_param_names() == [n for n, _ in ORDER[which]]
    shapes = {n: tuple(p.shape) for n, p in m.named_parameters()}
    assert [shapes[n] for n, _ in ORDER[which]] == [tuple(s) for _, s in ORDER[which]]


@pytest.mark.parametrize("which", ["performer", "gpt2", "stage1"])
def test_fused_adam_state_dict_is_indexed_like_reference_adam(which):
    """a torch.optim.Adam state dict over parameters in the REFERENCE order resumes in FusedAdam and comes back out
    with every moment under the same index (reference optim/ep*_optim.pt, train.py:318-326 / 342-347)"""
    m = _models()[which]
    g = torch.Generator().manual_seed(1)
    sd = {"state": {}, "param_groups": [{"lr": 3e-4, "betas": (0.9, 0.999), "eps": 1e-8}]}
    for i, (name, shape) in enumerate(ORDER[which]):
        sd["state"][i] = {"step": torch.tensor(7.0), "exp_avg": torch.randn(shape, generator=g),
                          "exp_avg_sq": torch.rand(shape, generator=g)}
    opt = FusedAdam(m, 1e-3)
    opt.load_state_dict(sd)
    assert opt._step == 7 and opt.param_groups[0]["lr"] == 3e-4
    sl = m._slices()
    for i, (name, shape) in enumerate(ORDER[which]):
        off, n, shp = sl[name]
        assert torch.equal(opt._m[off:off + n].view(shp), sd["state"][i]["exp_avg"]), name
        assert torch.equal(opt._v[off:off + n].view(shp), sd["state"][i]["exp_avg_sq"]), name
    back = opt.state_dict()
    assert list(back["state"].keys()) == list(range(len(ORDER[which])))
    for i in range(len(ORDER[which])):
        assert torch.equal(back["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
    # a checkpoint of another architecture is refused instead of being broadcast into the wrong slot
    bad = {"state": {0: {"step": torch.tensor(1.0), "exp_avg": torch.zeros(3), "exp_avg_sq": torch.zeros(3)}},
           "param_groups": [{}]}
    with pytest.raises(ValueError):
        opt.load_state_dict(bad)


# ------------------------------------------------------------------------------------------------
# data-parallel epoch order: same permutation on every rank, same number of steps
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,bs,world", [(103, 4, 2), (64, 4, 8), (10, 4, 4), (7, 2, 3)])
def test_epoch_batches_shard_one_permutation_evenly(n, bs, world):
    import random
    from emo_disentanger_b200.data.token_store import epoch_batches
    per_rank = []
    for r in range(world):
        random.seed(1000 + r)                   # the process-global state differs per rank: it must not matter
        per_rank.append(epoch_batches(n, bs, True, r, world, seed=5, epoch=3))
    assert len({len(b) for b in per_rank}) == 1                     # every rank takes the same number of steps
    flat = [i for b in per_rank for bi in b for i in bi]
    assert len(flat) == len(set(flat))                               # no piece twice
    nb = (n + bs - 1) // bs
    assert sum(len(b) for b in per_rank) == nb // world * world      # only the tail that does not fill a round is dropped
    # the interleaving of the ranks' batches is one cut of one permutation
    order = list(range(n))
    random.Random(5 + 3).shuffle(order)
    want = [order[i:i + bs] for i in range(0, n, bs)][:nb // world * world]
    got = [per_rank[i % world][i // world] for i in range(len(want))]
    assert got == want
    if per_rank[0]:
        assert epoch_batches(n, bs, True, 0, world, seed=5, epoch=4) != per_rank[0]     # a new epoch reshuffles
    # world 1 keeps every batch (reference DataLoader semantics)
    assert sum(len(b) for b in epoch_batches(n, bs, True, 0, 1, seed=1)) == n


def test_rank_strided_equal_steps_and_passthrough():
    from emo_disentanger_b200.scripts import common
    loader = list(range(11))
    got = [list(common.rank_strided(loader, r, 4)) for r in range(4)]
    assert got == [[0, 4], [1, 5], [2, 6], [3, 7]]                   # 11 // 4 * 4 = 8 batches, two steps per rank
    assert list(common.rank_strided(loader, 0, 1)) == loader

    class Sharded(list):
        rank_sharded = True
    assert list(common.rank_strided(Sharded([3, 1, 2]), 1, 2)) == [3, 1, 2]
    g0, g1 = common.shared_generator(7), common.shared_generator(7)
    assert torch.equal(torch.randperm(50, generator=g0), torch.randperm(50, generator=g1))
