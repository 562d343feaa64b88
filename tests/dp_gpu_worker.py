"""Worker of tests/test_dp_gpu.py (launched under torch.distributed.run, one rank per GPU, NCCL): the data-parallel
step over N GPUs at per-GPU batch b must equal the 1-GPU step at batch N*b -- same mean CE over the global non-pad
count, same all-reduced gradient, same clipped Adam update (reference stage2_accompaniment/train.py:71-81)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(dtype):
    from emo_disentanger_b200.stage2 import MusicPerformer
    from helpers import load_seeded
    from oracle import performer_oracle as PO
    V, L = 60, 2
    m = MusicPerformer(V, L, 8, 512, 2048, 512, dropout=0.0, use_segment_emb=True, n_segment_types=2,
                       favor_feature_dims=128, compute_dtype=dtype)
    load_seeded(m, PO.performer_state_shapes(V, L), 3)
    m = m.cuda().train()
    m.fixed_omegas = torch.randn(L, 64, 64, generator=torch.Generator().manual_seed(4)).cuda()
    return m, V


def main():
    from emo_disentanger_b200 import dp
    from emo_disentanger_b200.optim import FusedAdam
    rank, local, world = dp.init_from_env()
    out = {}
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        m, V = build(dtype)
        g = torch.Generator().manual_seed(7)
        B, T = 2 * world, 256
        tok = torch.randint(0, V - 1, (B, T), generator=g)
        seg = torch.randint(0, 2, (B, T), generator=g)
        tgt = torch.where(seg == 1, torch.roll(tok, -1, 1), torch.full_like(tok, V - 1))
        tgt[0, : T // 2] = V - 1                      # uneven non-pad counts per rank: the global count matters
        sync = dp.GradSync(m)
        sync.broadcast_params()
        opt = FusedAdam(m, lr=1e-3, max_grad_norm=0.5)          # the all-reduce is issued by hand to look at its result
        sl = slice(rank, B, world)
        acc = m.train_step(tok[sl].cuda(), seg[sl].cuda(), tgt[sl].cuda(), count_allreduce=sync.count_allreduce)
        acc = sync.allreduce_stats(acc.clone())
        sync.allreduce_grads()
        grad = m._flat_grad.clone()
        opt.step()
        torch.cuda.synchronize()
        if rank == 0:
            m1, _ = build(dtype)
            opt1 = FusedAdam(m1, lr=1e-3, max_grad_norm=0.5)
            acc1 = m1.train_step(tok.cuda(), seg.cuda(), tgt.cuda())
            grad1 = m1._flat_grad.clone()
            opt1.step()
            d = (m._flat - m1._flat).double()
            upd = (m1._flat.double() - build(dtype)[0]._flat.double())
            out[tag] = {"count": [float(acc[0]), float(acc1[0])], "loss_sum": [float(acc[1]), float(acc1[1])],
                        "grad_rms_rel": float((grad - grad1).double().pow(2).mean().sqrt() / grad1.double().pow(2).mean().sqrt()),
                        "param_diff_rel_to_update": float(d.pow(2).sum().sqrt() / upd.pow(2).sum().sqrt()),
                        "update_norm": float(upd.pow(2).sum().sqrt())}
        # every rank holds the same replica after the step
        chk = m._flat.double().sum().reshape(1)
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        if rank == 0:
            out[tag]["replica_checksums"] = [float(x) for x in lst]
    if rank == 0:
        json.dump(out, open(os.environ["EMO_DP_OUT"], "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
