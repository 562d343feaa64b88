"""GPU-resident token store for stage-2 training.

Drop-in for what `REMISkylineToMidiTransformerDataset` + `DataLoader` hand the reference loop
(stage2_accompaniment/dataloader.py:42-231, train.py:44-50,280-285): batches with the keys `id`, `dec_input`,
`dec_target`, `track_mask`, `length`, `chord_idx`, `melody_idx` -- already on the device.  The reference unpickles the
piece, converts event strings to ids, pads python lists and builds string-typed masks for EVERY item of EVERY epoch on
the host; here the pieces are tokenised once into one flat int32 array + offset tables in HBM, and a batch is one
kernel launch (csrc/dataset.cu).  The start bar of a sample is drawn on the host with `random.choice` over the same
admissible-bar list as the reference (:95-106,:112), so a seeded run picks the same windows."""
import ctypes as C
import pickle
import random

import numpy as np
import torch

from .. import _lib as L


def _p(t):
    return None if t is None else t.data_ptr()



def epoch_batches(n, batch_size, shuffle=True, rank=0, world=1, drop_last=False, seed=None, epoch=0):
    """The index lists of one epoch for one rank.  Data parallel needs every rank to cut THE SAME permutation and to
    take THE SAME number of steps (each optimiser step is a collective): with world > 1 the order comes from
    random.Random(seed + epoch) -- never the process-global `random` state, which differs per rank -- and the batch
    list is truncated to a multiple of `world`.  world == 1 with seed None keeps the reference behaviour (global
    `random`, every batch)."""
    order = list(range(n))
    if shuffle:
        if seed is None and world > 1:
            seed = 0
        (random if seed is None else random.Random(seed + epoch)).shuffle(order)
    batches = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    if drop_last and batches and len(batches[-1]) < batch_size:
        batches.pop()
    if world > 1:
        batches = batches[:len(batches) // world * world]
    return batches[rank::world]


class Stage2TokenStore:
    def __init__(self, pieces, event2idx, idx2event, model_dec_seqlen=10240, predict_key=False, device="cuda",
                 piece_ids=None):
        """pieces: iterable of (melody_pos, chord_pos, events) in the reference's pickle layout; events are
        'Name_Value' strings or {'name', 'value'} dicts (dataloader.py:29-39)."""
        self.event2idx, self.idx2event = event2idx, dict(idx2event)
        self.bar_token = event2idx['Bar_None']                                   # read_vocab, :66-74
        self.eos_token = event2idx['EOS_None']
        self.pad_token = len(event2idx)
        self.vocab_size = self.pad_token + 1
        self.model_dec_seqlen = int(model_dec_seqlen)
        self.predict_key = bool(predict_key)
        self.device = torch.device(device)
        toks, piece_off, bar_off, mel, c0, c1 = [], [0], [0], [], [], []
        self.piece_admissible_stbars = []
        for melody_pos, chord_pos, events in pieces:
            assert len(melody_pos) == len(chord_pos)
            ids = [event2idx['{}_{}'.format(e['name'], e['value'])] if isinstance(e, dict) else event2idx[e] for e in events]
            toks.extend(ids)
            piece_off.append(len(toks))
            mel.extend(int(m[0]) for m in melody_pos)
            c0.extend(int(c[0]) for c in chord_pos)
            c1.extend(int(c[1]) for c in chord_pos)
            bar_off.append(len(mel))
            n = len(ids)
            if n <= self.model_dec_seqlen:                                       # build_dataset, :95-106
                adm = [0]
            else:
                adm = []
                for bar in range(len(melody_pos)):
                    if n - melody_pos[bar][0] >= 0.5 * self.model_dec_seqlen:
                        adm.append(bar)
                    else:
                        break
            self.piece_admissible_stbars.append(adm)
        self.piece_ids = list(piece_ids) if piece_ids is not None else ["%d" % i for i in range(len(piece_off) - 1)]
        flags = np.zeros(self.vocab_size, dtype=np.uint8)
        for i in range(self.pad_token):                                          # :206-213 (PAD = 'Pad_None': neither)
            t = self.idx2event[i].split('_')[0]
            flags[i] = (1 if t == 'Chord' else 0) | (2 if t == 'Note' else 0)
        dev = self.device
        self.tokens = torch.tensor(np.asarray(toks, dtype=np.int32), device=dev)
        self.piece_off = torch.tensor(piece_off, dtype=torch.int64, device=dev)
        self.bar_off = torch.tensor(bar_off, dtype=torch.int64, device=dev)
        self.mel_start = torch.tensor(np.asarray(mel, dtype=np.int32), device=dev)
        self.ch_start = torch.tensor(np.asarray(c0, dtype=np.int32), device=dev)
        self.ch_end = torch.tensor(np.asarray(c1, dtype=np.int32), device=dev)
        self.flags = torch.tensor(flags, device=dev)
        self._sel_host = self._sel_dev = self._sel_ev = None

    @classmethod
    def from_files(cls, piece_files, vocab_file, **kw):
        """the reference's on-disk layout: dictionary.pkl = (event2idx, idx2event), one pickle per piece"""
        event2idx, idx2event = pickle.load(open(vocab_file, 'rb'))[:2]
        pieces = (pickle.load(open(f, 'rb'))[:3] for f in sorted(piece_files))
        ids = [f.split('/')[-1].replace('.pkl', '') for f in sorted(piece_files)]
        return cls(pieces, event2idx, idx2event, piece_ids=ids, **kw)

    def __len__(self):
        return self.piece_off.numel() - 1

    def batch(self, piece_idx, st_bars=None):
        """piece_idx: list of piece indices; st_bars: start bars (default: random.choice over the admissible bars,
        as get_sample_from_file does, :108-115).  Returns the reference batch dict, tensors on the device."""
        if not self.tokens.is_cuda:
            raise L.EmoError("Stage2TokenStore assembles batches on the GPU: build it with device='cuda'")
        B, T, dev = len(piece_idx), self.model_dec_seqlen, self.device
        if st_bars is None:
            st_bars = [random.choice(self.piece_admissible_stbars[i]) for i in piece_idx]
        if self._sel_host is None or self._sel_host.shape[1] < B:
            self._sel_host = torch.zeros(2, max(B, 64), dtype=torch.int32).pin_memory()
            self._sel_dev = torch.zeros(2, max(B, 64), dtype=torch.int32, device=dev)
            self._sel_ev = None
        if self._sel_ev is not None:
            self._sel_ev.synchronize()                    # the previous batch's copy has read the pinned buffer
        self._sel_host[:, :B] = torch.from_numpy(np.asarray([piece_idx, st_bars], dtype=np.int32))
        self._sel_dev.copy_(self._sel_host, non_blocking=True)
        self._sel_ev = torch.cuda.Event()
        self._sel_ev.record()
        sel = self._sel_dev
        out = torch.empty(5, B, T, dtype=torch.int64, device=dev)
        length = torch.empty(B, dtype=torch.int64, device=dev)
        L.check(L.lib().emo_stage2_batch(_p(self.tokens), _p(self.piece_off), _p(self.bar_off), _p(self.mel_start),
                                         _p(self.ch_start), _p(self.ch_end), _p(self.flags), _p(sel[0]), _p(sel[1]),
                                         _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _p(out[4]), _p(length), B, T,
                                         self.pad_token, self.eos_token, 1 if self.predict_key else 0,
                                         torch.cuda.current_stream().cuda_stream), "emo_stage2_batch")
        self.last_launch = (sel, out, length)
        return {'id': torch.tensor(list(piece_idx)), 'piece_id': [self.piece_ids[i] for i in piece_idx],
                'dec_input': out[0], 'dec_target': out[1], 'chords_mhot': 0, 'track_mask': out[2], 'length': length,
                'chord_idx': out[3], 'melody_idx': out[4]}

    def loader(self, batch_size, shuffle=True, rank=0, world=1, drop_last=False, seed=None, epoch=0):
        """epoch iterator with the DataLoader(shuffle=True) semantics of train.py:280-285; under data parallel
        every rank takes the batches rank::world of the same shuffled order (see epoch_batches)."""
        for bi in epoch_batches(len(self), batch_size, shuffle, rank, world, drop_last, seed, epoch):
            yield self.batch(bi)


class Stage1TokenStore:
    """GPU-resident stage-1 (lead-sheet) dataset: what `SkylineFullSongTransformerDataset` +
    `collate_fn` hand the reference loop (stage1_compose/dataloader.py:159-255,354-520 as train.py:230-262 builds it:
    first segment of every piece, no augmentation) -- `id`, `n_seg`, `dec_inp_0`, `dec_tgt_0`, `dec_seg_len_0`,
    `inp_chord_0`, `inp_melody_0` [B, model_dec_seqlen] int64, already on the device.  The reference's per-item work
    (unpickle, dict -> string -> id conversion, deepcopy, numpy padding and the never-read encoder features,
    :533-608) is replaced by one tokenisation pass and one kernel launch per batch (csrc/dataset.cu)."""

    def __init__(self, pieces, event2idx, idx2event, model_dec_seqlen=2400, model_max_bars=192, device="cuda",
                 piece_ids=None):
        """pieces: iterable of (bar_pos, events) in the reference's pickle layout"""
        self.event2idx, self.idx2event = event2idx, dict(idx2event)
        self.bar_token = event2idx['Bar_None']                                   # read_vocab, :343-352
        self.eos_token = event2idx['EOS_None']
        self.pad_token = len([k for k in event2idx if k != 'PAD_None'])
        self.vocab_size = self.pad_token + 1
        self.model_dec_seqlen, self.model_max_bars = int(model_dec_seqlen), int(model_max_bars)
        self.device = torch.device(device)
        toks, piece_off, seg_len = [], [0], []
        self.piece_bar_pos, self.piece_segments = [], []
        for bar_pos, events in pieces:
            ids = [event2idx['{}_{}'.format(e['name'], e['value'])] if isinstance(e, dict) else event2idx[e] for e in events]
            bp, n = list(bar_pos), len(ids)
            if bp[-1] == n:                                                      # build_dataset, :364-381
                bp = bp[:-1]
            if n - bp[-1] == 2:
                n = bp[-1]
                bp = bp[:-1]
            if len(bp) <= self.model_max_bars:
                bp.append(n - 1)
            else:
                bp = bp[:self.model_max_bars + 1]
            closing = self.eos_token if len(bp) <= self.model_max_bars else self.bar_token    # :431-435
            sample = ids[:bp[-1]] + [closing]
            end_bar = len(bp) - 1                                                # register_segments, :386-406
            for b in range(1, len(bp) - 1):
                if bp[b + 1] - bp[0] > self.model_dec_seqlen - 1:
                    end_bar = b
                    break
            E = bp[end_bar] - bp[0] + 1                                          # :483-491 (slices start at token 0)
            if E + 1 > len(sample):
                raise ValueError("piece %d: no event precedes the first bar -- the reference asserts on such a piece "
                                 "(dataloader.py:512)" % len(seg_len))
            toks.extend(sample)
            piece_off.append(len(toks))
            seg_len.append(E)
            self.piece_bar_pos.append(bp)
            self.piece_segments.append([(0, end_bar)])
        self.piece_ids = list(piece_ids) if piece_ids is not None else ["%d" % i for i in range(len(seg_len))]
        flags = np.zeros(self.vocab_size, dtype=np.uint8)
        for i in range(self.pad_token):                                          # :497-503 on ids
            t = self.idx2event[i].split('_')[0]
            flags[i] = (1 if t == 'Chord' else 0) | (2 if t == 'Note' else 0)
        dev = self.device
        self.tokens = torch.tensor(np.asarray(toks, dtype=np.int32), device=dev)
        self.piece_off = torch.tensor(piece_off, dtype=torch.int64, device=dev)
        self.seg_len = torch.tensor(np.asarray(seg_len, dtype=np.int32), device=dev)
        self.flags = torch.tensor(flags, device=dev)
        self._sel_host = self._sel_dev = self._sel_ev = None

    @classmethod
    def from_files(cls, piece_files, vocab_file, **kw):
        event2idx, idx2event = pickle.load(open(vocab_file, 'rb'))[:2]
        pieces = (pickle.load(open(f, 'rb'))[:2] for f in sorted(piece_files))
        ids = [f.split('/')[-1].replace('.pkl', '') for f in sorted(piece_files)]
        return cls(pieces, event2idx, idx2event, piece_ids=ids, **kw)

    def __len__(self):
        return self.piece_off.numel() - 1

    def batch(self, piece_idx):
        if not self.tokens.is_cuda:
            raise L.EmoError("Stage1TokenStore assembles batches on the GPU: build it with device='cuda'")
        B, T, dev = len(piece_idx), self.model_dec_seqlen, self.device
        if self._sel_host is None or self._sel_host.numel() < B:
            self._sel_host = torch.zeros(max(B, 64), dtype=torch.int32).pin_memory()
            self._sel_dev = torch.zeros(max(B, 64), dtype=torch.int32, device=dev)
            self._sel_ev = None
        if self._sel_ev is not None:
            self._sel_ev.synchronize()
        self._sel_host[:B] = torch.from_numpy(np.asarray(piece_idx, dtype=np.int32))
        self._sel_dev.copy_(self._sel_host, non_blocking=True)
        self._sel_ev = torch.cuda.Event()
        self._sel_ev.record()
        out = torch.empty(4, B, T, dtype=torch.int64, device=dev)
        length = torch.empty(B, dtype=torch.int64, device=dev)
        L.check(L.lib().emo_stage1_batch(_p(self.tokens), _p(self.piece_off), _p(self.seg_len), _p(self.flags),
                                         _p(self._sel_dev), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _p(length),
                                         B, T, self.pad_token, torch.cuda.current_stream().cuda_stream),
                "emo_stage1_batch")
        return {'id': torch.tensor(list(piece_idx)), 'piece_id': [self.piece_ids[i] for i in piece_idx],
                'st_seg': torch.zeros(B, dtype=torch.int64), 'n_seg': torch.ones(B, dtype=torch.int64),
                'dec_inp_0': out[0], 'dec_tgt_0': out[1], 'dec_seg_len_0': length, 'inp_chord_0': out[2],
                'inp_melody_0': out[3]}

    def loader(self, batch_size, shuffle=True, rank=0, world=1, drop_last=False, seed=None, epoch=0):
        """DataLoader(shuffle=True) semantics of stage1_compose/train.py:256-262; rank-strided under data parallel"""
        for bi in epoch_batches(len(self), batch_size, shuffle, rank, world, drop_last, seed, epoch):
            yield self.batch(bi)
