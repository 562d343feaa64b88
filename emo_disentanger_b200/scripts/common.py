"""Shared plumbing of the four scripts: pickle5 shim, reference-tree imports for the OUT-OF-SCOPE host
components (datasets, event->MIDI conversion; SURVEY 2: C8-C12), synthetic data, rank-0 logging."""
import os
import pickle
import sys
import torch

from ..synth import synthetic_batch


def install_pickle5_shim():
    """the reference imports `pickle5` (utils.py:2), which is python<3.8 only"""
    sys.modules.setdefault("pickle5", pickle)


def reference_module(stage_dir, name):
    """import `name` (dataloader / convert2midi / utils) from the reference tree the script is run in
    (cwd = reference repo root, as the reference README does) or from $EMO_REFERENCE_ROOT"""
    install_pickle5_shim()
    roots = [os.environ.get("EMO_REFERENCE_ROOT", ""), os.getcwd()]
    for r in roots:
        p = os.path.join(r, stage_dir)
        if r and os.path.isfile(os.path.join(p, name + ".py")):
            if p not in sys.path:
                sys.path.insert(0, p)
            return __import__(name)
    raise FileNotFoundError("%s/%s.py not found: run from the reference repo root (or set EMO_REFERENCE_ROOT), "
                            "or use a config with a `synthetic` data section" % (stage_dir, name))


class SyntheticStage2:
    """iterable of reference-shaped stage-2 batches"""

    def __init__(self, V, B, T, n, seed):
        self.vocab_size, self.pad_token = V, V - 1
        self.batches = []
        for i in range(n):
            tok, seg, tgt = synthetic_batch(V, B, T, seed + i)
            z = torch.zeros_like(tok)
            self.batches.append({"dec_input": tok, "dec_target": tgt, "track_mask": seg, "length": [T] * B,
                                 "chord_idx": z, "melody_idx": z})

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


class StoreEpochs:
    """a fresh pass over a GPU token store per `for batch in loader`: rank-sharded inside the store (no batch is
    assembled only to be discarded), the epoch's permutation drawn from Random(seed + epoch) so that every rank cuts
    the same order, the batch count truncated to a multiple of the world size"""
    rank_sharded = True

    def __init__(self, store, batch_size, shuffle=True, rank=0, world=1, seed=0):
        self.store, self.bs, self.shuffle, self.rank, self.world, self.seed = store, batch_size, shuffle, rank, world, seed
        self.epoch = 0

    def __len__(self):
        n = (len(self.store) + self.bs - 1) // self.bs
        return n // self.world if self.world > 1 else n

    def __iter__(self):
        ep, self.epoch = self.epoch, self.epoch + 1
        return self.store.loader(self.bs, shuffle=self.shuffle, rank=self.rank, world=self.world,
                                 seed=self.seed if (self.world > 1 or self.seed) else None, epoch=ep)


def rank_strided(loader, rank, world):
    """data parallel: rank r takes batches r, r+world, ... of an order every rank iterates alike; the
    tail that does not fill a round of `world` batches is dropped so that all ranks take the same number of
    optimiser steps (each one is a collective).  Loaders that shard themselves (StoreEpochs) pass through."""
    if getattr(loader, "rank_sharded", False) or world == 1:
        yield from loader
        return
    n = len(loader) // world * world
    for i, b in enumerate(loader):
        if i >= n:
            break
        if i % world == rank:
            yield b


def shared_generator(seed=0):
    """a torch.Generator for DataLoader(shuffle=True, generator=...): seeded alike on every rank and advanced only by
    the train loader, so rank 0's extra validation passes do not shift its training permutation"""
    g = torch.Generator()
    g.manual_seed(seed)
    return g
