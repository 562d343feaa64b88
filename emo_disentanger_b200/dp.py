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
    """Hooks for Stage2Base.train_step / FusedAdam.step.

    overlap=True (default): the gradient all-reduce is bucketed per layer and launched from the backward itself -- the
    model calls back after the last kernel that writes layer l's gradients (FlatModule._layer_done) and that layer's
    slice of the flat fp32 buffer is reduced on NCCL's stream while the backward of layer l - 1 runs; allreduce_grads()
    then reduces what is left (embeddings, output head, shared biases) and joins.  With gradient accumulation call
    `begin_step(last_micro_batch)`: only the last micro-batch may reduce early (the buffer still accumulates)."""

    def __init__(self, model, group=None, overlap=True):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1 and os.environ.get("EMO_DP_OVERLAP", "1") != "0"    # A/B switch
        self._work = []
        self._ranges = []
        self._comm = None
        if self.overlap:
            model.grad_hook = self._on_layer_done
        self._armed = True

    def begin_step(self, last_micro_batch=True):
        """gradient accumulation: early per-layer reduction only during the LAST micro-batch of an optimiser step"""
        self._armed = bool(last_micro_batch)

    def count_allreduce(self, count):
        """global number of non-pad targets: every rank scales its gradient by 1/global_count, so the
        summed gradient is the gradient of the global mean loss."""
        if self.world > 1:
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=self.group)
        return count

    def _on_layer_done(self, l):
        if not self._armed:
            return
        lo, hi = self.model.layer_grad_range(l)
        g = self.model._flat_grad
        if g.is_cuda:
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=g.device)
            ev = torch.cuda.Event()
            ev.record()                                   # everything that wrote this layer's gradients is before it
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ev)
                w = dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            w = dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._work.append(w)
        self._ranges.append((lo, hi))

    def allreduce_grads(self):
        if self.world <= 1:
            return
        g = self.model._flat_grad
        pos = 0
        for lo, hi in sorted(self._ranges) + [(g.numel(), g.numel())]:      # whatever no layer bucket covered
            if lo > pos:
                dist.all_reduce(g[pos:lo], op=dist.ReduceOp.SUM, group=self.group)
            pos = max(pos, hi)
        for w in self._work:
            w.wait()                                      # the current stream waits for the early buckets
        self._work, self._ranges = [], []

    def allreduce_stats(self, acc):
        """acc = [count(global already), loss_sum(local), n_correct(local)] -> global sums"""
        if self.world > 1:
            dist.all_reduce(acc[1:], op=dist.ReduceOp.SUM, group=self.group)
        return acc

    def broadcast_params(self, src=0):
        if self.world > 1:
            dist.broadcast(self.model._flat, src, group=self.group)
            self.model._lp_version = -1
