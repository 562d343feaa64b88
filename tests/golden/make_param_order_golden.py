"""Reference `named_parameters()` order and shapes of the three models (torch.optim state dicts are indexed by it),
taken from the UNMODIFIED reference modules imported from /root/reference (Performer over the fast_transformers
stand-in).  Writes tests/golden/param_order.json.  Run in the build container:  python tests/golden/make_param_order_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_import as R  # noqa: E402

kw = dict(n_token=20, n_layer=2, n_head=8, d_model=512, d_ff=2048, d_embed=512, use_segment_emb=True, n_segment_types=2)
out = {}
m = R.stage2_performer().MusicPerformer(favor_feature_dims=128, **kw)
out["performer"] = [[n, list(p.shape)] for n, p in m.named_parameters()]
m = R.stage2_gpt2().MusicGPT2(**kw)
out["gpt2"] = [[n, list(p.shape)] for n, p in m.named_parameters()]
m = R.stage1_model().PlainTransformer(512, 20, 2, 8, 512, 2048, 0, 64, pad_index=19, pre_lnorm=True)
out["stage1"] = [[n, list(p.shape)] for n, p in m.named_parameters()]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "param_order.json"), "w"), indent=0)
print({k: len(v) for k, v in out.items()})
