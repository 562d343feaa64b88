"""Mint tests/golden/stage1_dataset.npz: synthetic pieces in the reference's stage-1 pickle layout `(bar_pos, events)`
run through the UNMODIFIED `SkylineFullSongTransformerDataset` + its `collate_fn` (imported from /root/reference; build
container only) with the arguments of stage1_compose/train.py:230-241, next to the oracle restatement -- the script
asserts they are identical before writing.

    python tests/golden/make_stage1_dataset_golden.py"""
import contextlib, io, os, pickle, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dataset_oracle as DO
from emo_disentanger_b200.data import formats as F

REF = os.environ.get("EMO_REFERENCE", "/root/reference")
sys.modules.setdefault("pickle5", pickle)
sys.path.insert(0, os.path.join(REF, "stage1_compose"))
import dataloader as ref_dl                                     # noqa: E402  (the reference module, unmodified)

rng = np.random.RandomState(11)
spec = [(3, "eos"), (9, "eos"), (14, "empty_bar"), (6, "marker"), (1, "eos"), (11, "marker"), (2, "empty_bar"), (20, "eos")]
pieces = [DO.synthetic_stage1_piece(rng, nb, tail=tail) for nb, tail in spec]
pieces.append(DO.synthetic_stage1_piece(rng, 4, bar_len=(60, 80), tail="eos"))       # first bar alone longer than 48
e2i, i2e = F.build_dictionary([ev for _, ev in pieces], relative=True, **F.VOCAB_FLAGS["stage1_lead_sheet"])
out = {"n_pieces": len(pieces), "vocab": np.array([i2e[i] for i in range(len(i2e))])}
for p, (bp, ev) in enumerate(pieces):
    out["p%d_tokens" % p] = np.array([e2i[F.event_name(e)] for e in ev], dtype=np.int64)
    out["p%d_bar_pos" % p] = np.array(bp, dtype=np.int64)

with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "tests", "golden")) as tmp:
    F.save_dictionary(os.path.join(tmp, "dictionary.pkl"), e2i, i2e)
    os.makedirs(os.path.join(tmp, "events"))
    names = []
    for p, (bp, ev) in enumerate(pieces):
        names.append("p%02d.pkl" % p)
        F.save_piece(os.path.join(tmp, "events", names[-1]), list(bp), ev)
    configs = [(64, 192), (48, 5), (512, 192)]                     # (model_dec_seqlen, model_max_bars)
    out["configs"] = np.array(configs)
    for ci, (seqlen, max_bars) in enumerate(configs):
        with contextlib.redirect_stdout(io.StringIO()):
            ds = ref_dl.SkylineFullSongTransformerDataset(
                os.path.join(tmp, "events"), os.path.join(tmp, "dictionary.pkl"), pieces=names, do_augment=False,
                model_dec_seqlen=seqlen, model_max_bars=max_bars, max_n_seg=1, max_pitch=108, min_pitch=21,
                convert_dict_event=True)
            batch = ds.collate_fn([ds[i] for i in range(len(ds))])
        assert int(max(batch["n_seg"])) == 1
        is_chord, is_note = DO.vocab_flags(ds.idx2event, ds.pad_token)
        for key in ("dec_inp_0", "dec_tgt_0", "inp_chord_0", "inp_melody_0", "dec_seg_len_0"):
            out["c%d_%s" % (ci, key)] = np.asarray(batch[key]).astype(np.int64)
        for p in range(len(pieces)):
            o = DO.stage1_assemble(out["p%d_tokens" % p].tolist(), out["p%d_bar_pos" % p].tolist(), seqlen, max_bars,
                                   ds.pad_token, ds.eos_token, ds.bar_token, is_chord, is_note)
            for k in ("dec_inp", "dec_tgt", "inp_chord", "inp_melody"):
                assert np.array_equal(o[k], out["c%d_%s_0" % (ci, k)][p]), (ci, p, k)
            assert o["dec_seg_len"] == int(out["c%d_dec_seg_len_0" % ci][p]), (ci, p)
        print("config", (seqlen, max_bars), "seg lens", out["c%d_dec_seg_len_0" % ci].tolist())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stage1_dataset.npz"), **out)
print("wrote stage1_dataset.npz")
