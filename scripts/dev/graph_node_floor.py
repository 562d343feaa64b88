"""latency floor of a chain of dependent kernels in one CUDA graph: 62 x embed_rows (one CTA of 128 threads, a few hundred
bytes: the cheapest kernel of the decode step), with and without programmatic dependent launch -- what a decode step costs
before it does any work"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from emo_disentanger_b200 import ops, _lib
d, V, N = 512, 329, 62
tok = torch.zeros(1, dtype=torch.int64, device="cuda"); et = torch.randn(V, d, device="cuda")
outs = [torch.empty(1, d, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
for pdl in (0, 1):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.embed_rows(tok, None, None, et, None, None, outs[0], 1.0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    _lib.lib().emo_set_pdl(pdl)
    try:
        with torch.cuda.graph(g):
            for i in range(N):
                ops.embed_rows(tok, None, None, et, None, None, outs[i & 1], 1.0)
    finally:
        _lib.lib().emo_set_pdl(0)
    for _ in range(20): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print("pdl=%d: %d-node graph %.1f us per replay = %.2f us per node" % (pdl, N, us, us / N))
