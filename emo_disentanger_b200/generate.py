"""Decode loops: drop-ins for `generate_conditional` (reference stage2_accompaniment/inference.py:231-327)
and `generate_plain_xl` (stage1_compose/inference_utils.py:51-135).

Same arguments, same grammar rules (monotone Beat positions with the 256-retry abort, bar counting,
Track_LeadSheet interleave, PAD / EOS handling, first-token Key rule), same return values.  What changed
is where the arithmetic runs: the model step is incremental (decode.py / PlainTransformer.generate) and
`temperature` + `nucleus` run in the fused device sampler (sampler.cu) -- one int64 per token crosses to
the host instead of the whole logits row.  The uniform that drives the draw still comes from numpy's
global MT19937 stream (one `random_sample()` per draw, exactly what `np.random.choice(..., p=)` consumes),
so a seeded run follows the reference's stream.  A rejected sample (Beat going backwards, PAD, early EOS)
re-draws from the SAME logits: with a fixed feature map the model output for an unchanged prefix is
unchanged, so the reference's re-run of the model is skipped (stage 1 keeps the reference's quirk of
advancing the memory on a rejected step, SURVEY App. B.9, because it changes later logits).
"""
import time
import numpy as np
import torch

from . import ops
from .decode import Stage2Decoder, Stage1Decoder

MAX_DEC_INP_LEN = 2048          # stage2_accompaniment/inference.py:19


class DeviceSampler:
    """temperature + nucleus on the device (K11).  greedy=True -> argmax (bit-exact decode mode)."""

    def __init__(self, device, rows=1):
        self.rows = rows
        self.u = torch.zeros(rows, dtype=torch.float32, device=device)
        self._buf = torch.zeros(2 * rows, dtype=torch.int64, device=device)     # sampled ids | status words
        self.out = self._buf[:rows]
        self.status = self._buf[rows:].view(torch.int32)[:rows]
        self._u_host = torch.zeros(rows, dtype=torch.float32).pin_memory()
        self._host = torch.zeros(2 * rows, dtype=torch.int64).pin_memory()

    def draw(self, logits, V, temp, top_p, greedy=False, rng=None, banned=None):
        """logits fp32 [rows, >=V] (device).  Returns python ints (tokens); raises IndexError where the
        reference would (exactly one index above top_p, inference.py:93).  banned: uint8 [rows, V] device mask of
        inadmissible tokens (grammar-constrained draw); a row whose candidates are all banned returns -1."""
        rows = logits.shape[0]
        if not greedy:
            r = np.random if rng is None else rng
            for i in range(rows):
                self._u_host[i] = r.random_sample()
            self.u.copy_(self._u_host, non_blocking=True)
        ops.sample(logits, V, temp, top_p, self.u, self.out, self.status, greedy=greedy, banned=banned)
        self._host.copy_(self._buf, non_blocking=True)                          # ONE small D2H into pinned memory
        torch.cuda.current_stream().synchronize()
        st = self._host[self.rows:].view(torch.int32)[:rows].tolist()
        if 1 in st:
            raise IndexError("index 1 is out of bounds for axis 0 with size 1")
        ids = self._host[:rows].tolist()
        return [(-1 if s_ == 2 else t) for t, s_ in zip(ids, st)]


def get_position_idx(event):
    return int(event.split('_')[-1])


# ------------------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------------------
def generate_conditional(model, event2idx, idx2event, lead_sheet_events, primer,
                         max_events=10000, skip_check=False, max_bars=None,
                         temp=1.2, top_p=0.9, inadmissibles=None,
                         model_type="performer", greedy=False, decoder=None, rng=None, verbose=True,
                         device_grammar=False):
    """device_grammar=True (opt-in): the rejection rules of the loop below (Beat positions must not
    go backwards, no PAD, no EOS before the last bar) are applied INSIDE the device sampler as a mask over the nucleus
    candidates -- the distribution the reference's reject-and-redraw loop samples from, without the wasted model calls.
    The numpy RNG stream is consumed differently (one uniform per accepted token), so seeded runs differ from the
    reference's; off by default."""
    if inadmissibles is not None:
        raise NotImplementedError("the reference always passes inadmissibles=None (inference.py:467)")
    say = print if verbose else (lambda *a, **k: None)
    generated = primer + [event2idx['Track_LeadSheet']] + lead_sheet_events[0] + [event2idx['Track_Full']]
    seg_inp = [0 for _ in range(len(generated))]
    seg_inp[-1] = 1

    target_bars, generated_bars = len(lead_sheet_events), 0
    if max_bars is not None:
        target_bars = min(max_bars, target_bars)

    dec = decoder if decoder is not None else Stage2Decoder(model, batch=1, max_len=MAX_DEC_INP_LEN)
    dec.reset(0)
    sampler = DeviceSampler(dec.dev)
    V = model.n_token
    fed = 0                       # tokens of `generated` already folded into the decode state
    logits = None
    predrawn = None               # token drawn inside the fused step graph, not yet consumed by the rules below
    banned = banned_host = None
    if device_grammar and not skip_check and not greedy:
        beat_pos = {i: get_position_idx(e) for i, e in idx2event.items() if 'Beat' in e}
        pad_id, eos_id = event2idx.get('PAD_None', V - 1), event2idx['EOS_None']
        banned_host = torch.zeros(1, V, dtype=torch.uint8).pin_memory()
        banned = torch.zeros(1, V, dtype=torch.uint8, device=dec.dev)
        ban_state = None

        def update_banned(cur_pos_, last_bar_):
            nonlocal ban_state
            if ban_state == (cur_pos_, last_bar_):
                return
            ban_state = (cur_pos_, last_bar_)
            banned_host.zero_()
            for i, ppos in beat_pos.items():
                if ppos < cur_pos_:
                    banned_host[0, i] = 1
            banned_host[0, pad_id] = 1
            if not last_bar_:
                banned_host[0, eos_id] = 1
            banned.copy_(banned_host, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the pinned buffer is rewritten on the next change

    steps = 0
    time_st = time.time()
    cur_pos = 0
    failed_cnt = 0

    while generated_bars < target_bars:
        assert len(generated) == len(seg_inp)
        if banned is not None:
            update_banned(cur_pos, generated_bars >= target_bars - 1)
        if len(generated) < MAX_DEC_INP_LEN:
            if fed < len(generated):          # fold the not-yet-seen suffix (primer, new token, lead-sheet bar)
                n_new = len(generated) - fed
                if n_new == 1 and fed > 0 and dec.use_graph:
                    # the common case: one new token -> model step and draw fused in one CUDA graph
                    u = 0.0 if greedy else (np.random if rng is None else rng).random_sample()
                    ids, st = dec.step_sample([generated[-1]], [seg_inp[-1]], [u], temp, top_p, greedy=greedy, banned=banned)
                    if st[0] == 1:
                        raise IndexError("index 1 is out of bounds for axis 0 with size 1")
                    if st[0] == 2:
                        ids = [-1]
                    logits = dec.logits[0:1, :V]
                    predrawn = ids[0]
                elif n_new == 1 and fed > 0:
                    logits = dec.step([generated[-1]], [seg_inp[-1]])[0:1]
                else:
                    logits = dec.append(0, generated[fed:], seg_inp[fed:])[None]
                fed = len(generated)
        else:                                 # window slides -> absolute positions shift -> full recompute
            dev = dec.dev
            dec_input = torch.tensor([generated[-MAX_DEC_INP_LEN:]]).long().to(dev)
            dec_seg_inp = torch.tensor([seg_inp[-MAX_DEC_INP_LEN:]]).long().to(dev)
            prev_omegas = getattr(model, "fixed_omegas", None)
            if dec.is_performer:
                model.fixed_omegas = dec.omegas
            try:
                with torch.no_grad():
                    logits = model(dec_input, seg_inp=dec_seg_inp, keep_last_only=True).float().contiguous()
            finally:
                if dec.is_performer:        # later forward / train_step calls redraw the feature map again
                    model.fixed_omegas = prev_omegas

        if predrawn is not None:
            word, predrawn = predrawn, None
        else:
            word = sampler.draw(logits, V, temp, top_p, greedy=greedy, rng=rng, banned=banned)[0]
        if word < 0:                  # every nucleus candidate is inadmissible: the reference would retry 256 times
            say('[FATAL] model stuck, exiting with generated events ...')
            return generated
        word_event = idx2event[word]

        if not skip_check:
            if 'Beat' in word_event:
                event_pos = get_position_idx(word_event)
                if not event_pos >= cur_pos:
                    failed_cnt += 1
                    say('[info] position not increasing, failed cnt:', failed_cnt)
                    if failed_cnt >= 256 or greedy:
                        say('[FATAL] model stuck, exiting with generated events ...')
                        return generated
                    continue
                else:
                    cur_pos = event_pos
                    failed_cnt = 0

        if word_event == 'Track_LeadSheet':
            steps += 1
            generated.append(word)
            seg_inp.append(0)
            generated_bars += 1
            say('[info] generated {} bars, #events = {}'.format(generated_bars, len(generated)))

            if generated_bars < target_bars:
                generated.extend(lead_sheet_events[generated_bars])
                seg_inp.extend([0 for _ in range(len(lead_sheet_events[generated_bars]))])

                generated.append(event2idx['Track_Full'])
                seg_inp.append(1)
                cur_pos = 0
            continue

        if word_event == 'PAD_None' or (word_event == 'EOS_None' and generated_bars < target_bars - 1):
            if greedy:
                say('[FATAL] greedy decode stuck on an inadmissible token')
                return generated
            continue
        elif word_event == 'EOS_None' and generated_bars == target_bars - 1:
            say('[info] gotten eos')
            generated.append(word)
            break

        generated.append(word)
        seg_inp.append(1)
        steps += 1

        if len(generated) > max_events:
            say('[info] max events reached')
            break

    say('-- generated events:', len(generated))
    say('-- time elapsed  : {:.2f} secs'.format(time.time() - time_st))
    say('-- time per event: {:.2f} secs'.format((time.time() - time_st) / len(generated)))
    return generated[:-1]


def _conditional_rules(event2idx, idx2event, lead_sheet_events, primer, max_events, skip_check, max_bars, greedy, say):
    """The bookkeeping and rejection rules of generate_conditional (stage2_accompaniment/inference.py:231-327) as a
    coroutine, so that several sequences can share one batched model step: yields ('step', tokens, segs) when the tokens
    not yet seen by the model must be folded in and a token drawn from the resulting logits, ('redraw',) when the last
    draw was rejected and another one is wanted from the SAME logits; is sent the drawn token id (-1 = every candidate
    inadmissible).  Returns the generated list."""
    generated = primer + [event2idx['Track_LeadSheet']] + lead_sheet_events[0] + [event2idx['Track_Full']]
    seg_inp = [0 for _ in range(len(generated))]
    seg_inp[-1] = 1
    target_bars, generated_bars = len(lead_sheet_events), 0
    if max_bars is not None:
        target_bars = min(max_bars, target_bars)
    fed, cur_pos, failed_cnt = 0, 0, 0
    while generated_bars < target_bars:
        if fed < len(generated):
            word = yield ('step', generated[fed:], seg_inp[fed:])
            fed = len(generated)
        else:
            word = yield ('redraw',)
        if word < 0:
            say('[FATAL] model stuck, exiting with generated events ...')
            return generated
        word_event = idx2event[word]
        if not skip_check and 'Beat' in word_event:
            event_pos = get_position_idx(word_event)
            if not event_pos >= cur_pos:
                failed_cnt += 1
                if failed_cnt >= 256 or greedy:
                    say('[FATAL] model stuck, exiting with generated events ...')
                    return generated
                continue
            cur_pos, failed_cnt = event_pos, 0
        if word_event == 'Track_LeadSheet':
            generated.append(word)
            seg_inp.append(0)
            generated_bars += 1
            if generated_bars < target_bars:
                generated.extend(lead_sheet_events[generated_bars])
                seg_inp.extend([0 for _ in range(len(lead_sheet_events[generated_bars]))])
                generated.append(event2idx['Track_Full'])
                seg_inp.append(1)
                cur_pos = 0
            continue
        if word_event == 'PAD_None' or (word_event == 'EOS_None' and generated_bars < target_bars - 1):
            if greedy:
                say('[FATAL] greedy decode stuck on an inadmissible token')
                return generated
            continue
        elif word_event == 'EOS_None' and generated_bars == target_bars - 1:
            generated.append(word)
            break
        generated.append(word)
        seg_inp.append(1)
        if len(generated) > max_events:
            say('[info] max events reached')
            break
    return generated[:-1]


def generate_conditional_batch(model, event2idx, idx2event, lead_sheets, primers, temps, top_p=0.9, max_events=10000,
                               skip_check=False, max_bars=None, greedy=False, decoder=None, rng=None, verbose=True):
    """Several accompaniments decoded in LOCKSTEP -- e.g. the four emotion quadrants of a lead-sheet pair, which the
    reference generates one after the other (stage2_accompaniment/inference.py:433-468).  Every sequence keeps its own
    rule state (the same rules as generate_conditional); one ragged batched model step + device sampler (one temperature
    per row) serves all of them per iteration, a rejected draw is re-drawn from that row's logits, a lead-sheet bar
    appended mid-stream is folded into that row's state alone.  Greedy tokens equal generate_conditional's; sampled runs
    consume the numpy RNG in a different order than four sequential calls.  Returns the list of token lists."""
    say = print if verbose else (lambda *a, **k: None)
    n = len(lead_sheets)
    V = model.n_token
    dec = decoder if decoder is not None else Stage2Decoder(model, batch=n, max_len=MAX_DEC_INP_LEN)
    if dec.B != n:
        raise ValueError("decoder batch %d != number of sequences %d" % (dec.B, n))
    dec.reset()
    r = np.random if rng is None else rng
    sampler = DeviceSampler(dec.dev)
    # the per-row temperatures live in a buffer owned by the decoder: its ADDRESS is part of the captured step graph's
    # key, so a reused decoder replays the graph of the previous piece instead of capturing a new one
    if getattr(dec, "_t_rows", None) is None:
        dec._t_rows = torch.empty(dec.B, dtype=torch.float32, device=dec.dev)
    dec._t_rows.copy_(torch.tensor([float(t) for t in temps], dtype=torch.float32))
    t_rows = dec._t_rows
    gens = [_conditional_rules(event2idx, idx2event, lead_sheets[b], list(primers[b]), max_events, skip_check, max_bars, greedy, say)
            for b in range(n)]
    results = [None] * n
    req = [None] * n

    def advance(b, word):
        try:
            req[b] = gens[b].send(word) if word is not None else next(gens[b])
        except StopIteration as stop:
            results[b], req[b] = stop.value, None

    for b in range(n):
        advance(b, None)
    last_tok, last_seg = [0] * n, [1] * n
    while any(q is not None for q in req):
        # rejected draws first: another draw from the row's own logits, until every live sequence wants a model step
        for b in range(n):
            while req[b] is not None and req[b][0] == 'redraw':
                w = sampler.draw(dec.logits[b:b + 1, :V], V, float(temps[b]), top_p, greedy=greedy, rng=rng)[0]
                advance(b, w)
        live = [b for b in range(n) if req[b] is not None]
        if not live:
            break
        if max(dec.pos_host[b] + len(req[b][1]) for b in live) > dec.max_len:
            raise RuntimeError("sequence longer than the decode state (max_len=%d)" % dec.max_len)
        for b in live:                          # all but the last pending token of a row: that row's own block
            toks, segs = req[b][1], req[b][2]
            if len(toks) > 1:
                dec.append(b, toks[:-1], segs[:-1])
            last_tok[b], last_seg[b] = toks[-1], segs[-1]
        for b in range(n):                      # finished rows idle along: keep their position inside the state's range
            if req[b] is None and dec.pos_host[b] > dec.max_len - 8:
                dec.pos[b] = 0
                dec.pos_host[b] = 0
        us = [0.0 if greedy else r.random_sample() for _ in range(n)]
        ids, st = dec.step_sample(last_tok, last_seg, us, t_rows, top_p, greedy=greedy)
        for b in live:
            if st[b] == 1:
                raise IndexError("index 1 is out of bounds for axis 0 with size 1")
            advance(b, -1 if st[b] == 2 else ids[b])
    return results


# ------------------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------------------
MAJOR_KEY = np.array(['C', 'C#', 'D', 'D#', 'E', 'F', 'F#', 'G', 'G#', 'A', 'A#', 'B'])
MINOR_KEY = np.array(['c', 'c#', 'd', 'd#', 'e', 'f', 'f#', 'g', 'g#', 'a', 'a#', 'b'])


def match_emotion_key(emotion, key):
    # stage1_compose/inference_utils.py:138-143
    if emotion in ['Q1', 'Q4', 'Positive'] and key in MAJOR_KEY:
        return True
    if emotion in ['Q2', 'Q3', 'Negative'] and key in MINOR_KEY:
        return True
    return False


def generate_plain_xl(model, event2idx, idx2event, max_bars=160,
                      max_events=2048, primer=None, temp=1.2, top_p=0.9,
                      prompt_bars=None, representation='functional', key_determine=None,
                      greedy=False, rng=None, verbose=True, decoder=None, incremental=True):
    """incremental=True (default): the model call of the reference loop runs through Stage1Decoder (K | V cache, one
    CUDA graph per token, model step + draw fused) -- the same tokens as feeding the hidden-state memory back through
    PlainTransformer.generate, which incremental=False still does (inference_utils.py:95-104)."""
    say = print if verbose else (lambda *a, **k: None)
    if primer is None:
        generated = [event2idx['Bar_None']]
        target_bars, generated_bars = max_bars, 0
    else:
        generated = [event2idx[e] for e in primer]
        target_bars, generated_bars = max_bars, prompt_bars if prompt_bars is not None else 0

    device = next(model.parameters()).device
    sampler = DeviceSampler(device)
    V = model.vocab_size
    steps = 0
    time_st = time.time()
    cur_pos = 0
    failed_cnt = 0
    mems = tuple()
    dec = decoder
    if dec is None and incremental:
        dec = Stage1Decoder(model, batch=1, max_len=max_events + 1024)
    if dec is not None:
        dec.reset()
    while generated_bars < target_bars:
        first_key = representation in ['functional', 'key'] and len(generated) == 1
        t_, p_ = (1.1, 0.97) if first_key else (temp, top_p)
        predrawn = None
        if dec is not None:
            # the reference feeds the whole list while no token has been accepted, the last token afterwards -- also
            # after a rejected draw: its memory advances by a duplicate row, and so does the cache here
            feed = generated if steps == 0 else [generated[-1]]
            for tk in feed[:-1]:
                dec.step([tk])
            if dec.use_graph:                                   # model step and draw in one CUDA graph
                u = 0.0 if greedy else (np.random if rng is None else rng).random_sample()
                ids, st = dec.step_sample([feed[-1]], [u], t_, p_, greedy=greedy)
                if st[0] == 1:
                    raise IndexError("index 1 is out of bounds for axis 0 with size 1")
                predrawn = ids[0]
                logits = dec.logits[0:1, :V]
            else:
                logits = dec.step([feed[-1]])[0:1]
        else:
            if steps == 0:
                dec_input = torch.LongTensor([generated]).to(device)
                dec_input = dec_input.permute(1, 0) if len(generated) > 1 else dec_input
            else:
                dec_input = torch.LongTensor([[generated[-1]]]).to(device)
            logits, mems = model.generate(dec_input, mems)         # memory advances even if the draw is rejected
            logits = logits.float().view(1, -1)

        if first_key:
            word = predrawn if predrawn is not None else sampler.draw(logits, V, 1.1, 0.97, greedy=greedy, rng=rng)[0]
            if key_determine == 'rule':
                emotion_label = idx2event[generated[0]].split('_')[1]
                key_event = idx2event[word]
                if key_event.split('_')[0] != 'Key':
                    raise ValueError('[info] key generation failed')
                key_label = key_event.split('_')[1]
                if not match_emotion_key(emotion_label, key_label):
                    if greedy:
                        raise ValueError('[info] greedy key does not match the emotion')
                    continue
            word_event = idx2event[word]
        else:
            word = predrawn if predrawn is not None else sampler.draw(logits, V, temp, top_p, greedy=greedy, rng=rng)[0]
            word_event = idx2event[word]

        if 'Key' in word_event:
            say('[info] generated {}, #events = {}'.format(word_event, len(generated)))

        if 'Beat' in word_event:
            event_pos = get_position_idx(word_event)
            if not event_pos >= cur_pos:
                failed_cnt += 1
                say('[info] position not increasing, failed cnt:', failed_cnt)
                if failed_cnt >= 256 or greedy:
                    say('[FATAL] model stuck, exiting ...')
                    return None, time.time() - time_st
                continue
            else:
                cur_pos = event_pos
                failed_cnt = 0

        if 'Bar' in word_event:
            generated_bars += 1
            cur_pos = 0
            say('[info] generated {} bars, #events = {}'.format(generated_bars, len(generated)))
        if word_event == 'PAD_None':
            if greedy:
                return None, time.time() - time_st
            continue

        generated.append(word)
        steps += 1

        if len(generated) > max_events:
            say('[info] max events reached')
            break
        if word_event == 'EOS_None':
            say('[info] gotten eos')
            break

    say('-- generated events:', len(generated))
    say('-- time elapsed: {:.2f} secs'.format(time.time() - time_st))
    return generated[:-1], time.time() - time_st
