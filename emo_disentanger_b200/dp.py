"""Data-parallel plumbing: one process per GPU, full replica, ONE all-reduce(sum) of the
flat fp32 gradient buffer per optimiser step over NCCL (NVLink 5 / NVSwitch), plus the all-reduce of
the non-pad target count so that the mean-CE and the global-norm clip match the 1-GPU semantics of
reference stage2_accompaniment/train.py:71-81 at N x batch.  The reference has no multi-GPU path; world
size 1 degenerates to no communication at all."""
import os
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, local, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


class GradSync:
    """Hooks for Stage2Base.train_step / FusedAdam.step."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def count_allreduce(self, count):
        """global number of non-pad targets: every rank scales its gradient by 1/global_count, so the
        summed gradient is the gradient of the global mean loss."""
        if self.world > 1:
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=self.group)
        return count

    def allreduce_grads(self):
        if self.world > 1:
            dist.all_reduce(self.model._flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def allreduce_stats(self, acc):
        """acc = [count(global already), loss_sum(local), n_correct(local)] -> global sums"""
        if self.world > 1:
            dist.all_reduce(acc[1:], op=dist.ReduceOp.SUM, group=self.group)
        return acc

    def broadcast_params(self, src=0):
        if self.world > 1:
            dist.broadcast(self.model._flat, src, group=self.group)
            self.model._lp_version = -1
