"""Mint tests/golden/dataset_small.npz: synthetic pieces in the reference's pickle format, run through the UNMODIFIED
reference `REMISkylineToMidiTransformerDataset` (imported from /root/reference; build container only), together with
the oracle restatement's outputs for the same (piece, start bar) picks -- the script asserts they are identical before
writing.  Fixture: flattened pieces + the reference outputs.

    python tests/golden/make_dataset_golden.py"""
import os, pickle, random, sys, tempfile, types
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dataset_oracle as DO
from emo_disentanger_b200.synth import synthetic_vocab

REF = os.environ.get("EMO_REFERENCE", "/root/reference")
sys.modules.setdefault("pickle5", pickle)
sys.path.insert(0, os.path.join(REF, "stage2_accompaniment"))
import dataloader as ref_dl                                     # noqa: E402  (the reference module, unmodified)

V, SEQ = 120, 96
e2i, i2e = synthetic_vocab(V, 2)
e2i = {k: v for k, v in e2i.items() if k != 'PAD_None'}          # dictionary.pkl has no PAD entry: pad = len(vocab)
i2e = {v: k for k, v in e2i.items()}
rng = np.random.RandomState(7)
tmp = tempfile.mkdtemp()
pickle.dump((e2i, i2e), open(os.path.join(tmp, "dictionary.pkl"), "wb"))
pieces = []
os.makedirs(os.path.join(tmp, "events"))
for p, nb in enumerate([2, 5, 9, 14, 1, 7]):                      # shorter and longer than SEQ, a one-bar piece
    mel, ch, ev = DO.synthetic_piece(e2i, nb, rng, as_dicts=(p % 2 == 1))
    pickle.dump((mel, ch, ev), open(os.path.join(tmp, "events", "p%02d.pkl" % p), "wb"))
    pieces.append((mel, ch, [e2i[DO.event_name(e)] for e in ev]))

out = {"V": V + 0, "seqlen": SEQ, "n_pieces": len(pieces)}
for p, (mel, ch, toks) in enumerate(pieces):
    out["p%d_tokens" % p] = np.array(toks, dtype=np.int64)
    out["p%d_mel" % p] = np.array(mel, dtype=np.int64)
    out["p%d_ch" % p] = np.array(ch, dtype=np.int64)
picks = []
for predict_key in (False, True):
    ds = ref_dl.REMISkylineToMidiTransformerDataset(os.path.join(tmp, "events"), os.path.join(tmp, "dictionary.pkl"),
                                                    model_dec_seqlen=SEQ, predict_key=predict_key)
    pad, eos = ds.pad_token, ds.eos_token
    is_chord, is_note = DO.vocab_flags(ds.idx2event, pad)
    for p in range(len(pieces)):
        adm = ds.piece_admissible_stbars[p]
        assert adm == DO.admissible_stbars(len(pieces[p][2]), pieces[p][0], SEQ), (p, adm)
        for st in adm:
            # the reference picks the start bar with random.choice: pin it
            ds.piece_admissible_stbars[p] = [st]
            ref = ds[p]
            ds.piece_admissible_stbars[p] = adm
            mine = DO.assemble(pieces[p][2], pieces[p][0], pieces[p][1], st, SEQ, pad, eos, is_chord, is_note, predict_key)
            for k in ("dec_input", "dec_target", "track_mask", "chord_idx", "melody_idx"):
                assert np.array_equal(np.asarray(ref[k]), mine[k]), (p, st, k)
            assert ref["length"] == mine["length"]
            i = len(picks)
            picks.append((p, st, int(predict_key)))
            for k in ("dec_input", "dec_target", "track_mask", "chord_idx", "melody_idx"):
                out["s%d_%s" % (i, k)] = np.asarray(ref[k]).astype(np.int64)
            out["s%d_length" % i] = ref["length"]
out["picks"] = np.array(picks, dtype=np.int64)
out["pad"], out["eos"] = pad, eos
out["is_chord"], out["is_note"] = is_chord, is_note
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataset_small.npz"), **out)
print("wrote dataset_small.npz: %d pieces, %d (piece, start bar, predict_key) samples -- reference == oracle" % (len(pieces), len(picks)))
