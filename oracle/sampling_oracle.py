"""CPU oracle: temperature / nucleus (top-p) sampling.  TEST INFRASTRUCTURE.

Restates reference stage2_accompaniment/inference.py:71-100 and
stage1_compose/inference_utils.py:14-41 (identical maths):
  * softmax(logits / t) in the logits' own dtype (fp32 from the model),
  * renormalise by a *sequential* sum, sort descending, cumulative sum,
  * cut at the SECOND index whose cumulative mass exceeds p (``where(..)[0][1]``; IndexError
    when exactly one index exceeds -- kept), fallback top-3 when none exceeds,
  * renormalise candidates in float64, draw with the legacy numpy ``choice`` rule
    (inverse CDF with one uniform: ``searchsorted(cdf, u, side='right')``).
``greedy`` = argmax (the bit-exact decode mode of the north star).
PINNED against the reference functions by oracle/validate_against_reference.py.
"""
import numpy as np


def temperature_probs(logits, t):
    logits = np.asarray(logits)
    z = np.exp(logits / t)
    probs = z / np.sum(z)
    if np.isnan(probs).any():          # reference retries in float128
        z = np.exp(logits.astype(np.longdouble) / t)
        probs = (z / np.sum(z)).astype(float)
    return probs


def nucleus_candidates(probs, p):
    """Returns (candidate indices, float64 candidate probabilities)."""
    probs = np.array(probs, copy=True)
    seq_sum = np.cumsum(probs, dtype=probs.dtype)[-1]       # == python sum(): sequential
    probs = probs / seq_sum
    order = np.argsort(probs)[::-1]
    csum = np.cumsum(probs[order])
    above = np.nonzero(csum > p)[0]
    if above.size > 0:
        last = above[1]                                      # IndexError if size == 1 (kept)
        cand = order[:last]
    else:
        cand = order[:3]
    cp = probs[cand].astype(np.float64)
    tot = 0.0
    for x in cp:                                             # sequential float64 sum
        tot += x
    return cand, cp / tot


def choose(cand, cand_probs, u):
    """numpy legacy RandomState.choice(cand, p=cand_probs) given its uniform draw ``u``."""
    cdf = np.cumsum(cand_probs)
    cdf /= cdf[-1]
    return cand[int(np.searchsorted(cdf, u, side="right"))]


def sample(logits, t, p, u):
    cand, cp = nucleus_candidates(temperature_probs(logits, t), p)
    return int(choose(cand, cp, u))


def greedy(logits):
    return int(np.argmax(np.asarray(logits)))
