// Stage-2 training batches assembled on the device from a GPU-resident token store.
//
// Replaces the per-item work of reference stage2_accompaniment/dataloader.py:178-231
// (REMISkylineToMidiTransformerDataset.__getitem__: pickle_load of the piece, event -> id conversion, python-list
// padding, make_target_and_mask :127-145 / make_target_and_mask_predict :147-173, the string-typed chord / melody
// masks :206-213) + the DataLoader collation + the H2D copies of train.py:44-50.  The pieces are tokenised ONCE into
// one flat int32 array with per-piece / per-bar offset tables; a batch is one launch (one CTA per sample) that writes
// dec_input, dec_target, track_mask, chord_idx, melody_idx [B, T] int64 and length [B] -- the tensors the reference
// loop feeds the model.  Integer work, bit-exact against the reference (tests/golden/dataset_small.npz).
#include "common.cuh"

namespace {

struct BatchArgs {
  const int32_t* tokens;      // all pieces, concatenated
  const int64_t* piece_off;   // [P + 1] into tokens
  const int64_t* bar_off;     // [P + 1] into the bar tables
  const int32_t* mel_start;   // per bar: start of the lead-sheet span (melody_pos[b][0])
  const int32_t* ch_start;    // per bar: Full-track span [ch_start, ch_end)   (chord_pos[b])
  const int32_t* ch_end;
  const uint8_t* flags;       // [V]: bit 0 = 'Chord_*' event, bit 1 = 'Note_*' event
  const int32_t* sel_piece;   // [B]
  const int32_t* sel_stbar;   // [B]
  int64_t *inp, *tgt, *mask, *chord_idx, *melody_idx, *length;
  int T, pad, eos, predict_key;
};

// One thread per output position (grid = T/256 x B): the bar whose Full-track span covers position t is found by a
// binary search over the (increasing, disjoint) span starts, so there is no per-bar loop and no barrier -- every
// position is written exactly once, 40 bytes per token of pure streaming stores.
__global__ void __launch_bounds__(256) stage2_batch_kernel(const BatchArgs a) {
  constexpr int SB = 1024;                               // bar spans staged in shared memory (pieces have ~100 bars)
  __shared__ int s_c0[SB], s_c1[SB];
  const int b = blockIdx.y, T = a.T;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = a.sel_piece[b], st = a.sel_stbar[b];
  const int64_t base = a.piece_off[p];
  const int plen = (int)(a.piece_off[p + 1] - base);
  const int64_t bo = a.bar_off[p];
  const int nb = (int)(a.bar_off[p + 1] - bo);
  const int hdr = a.mel_start[bo];                       // events before the first bar are always kept (:186)
  const int skip = a.mel_start[bo + st] - hdr;           // events dropped between the header and the start bar
  const int seqlen = plen - skip;                        // header + everything from the start bar on
  const bool staged = nb - st <= SB;
  if (staged)
    for (int i = threadIdx.x; i < nb - st; i += blockDim.x) {
      s_c0[i] = a.ch_start[bo + st + i] - skip;
      s_c1[i] = a.ch_end[bo + st + i] - skip;
    }
  __syncthreads();
  if (t >= T) return;
  // token at position u of the (untruncated) sample: PAD past its end (pad_sequence, :117-125)
  auto token = [&](int u) -> int64_t { return u < seqlen ? (int64_t)a.tokens[base + (u < hdr ? u : u + skip)] : (int64_t)a.pad; };

  int64_t tg = a.pad, mk = 0;
  if (a.predict_key) {                                   // :153-158, written before the bar spans in the reference
    if (t == 0) { mk = 2; tg = token(1); }
    if (t == 1) mk = 3;
  }
  // last bar in [st, nb) whose span starts at or before t (:131-143: spans of the bars from the start bar on)
  int lo = st, hi = nb;                                  // invariant: start(lo - 1) <= t < start(hi)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int c0 = staged ? s_c0[mid - st] : a.ch_start[bo + mid] - skip;
    if (c0 <= t) lo = mid + 1; else hi = mid;
  }
  const int bar = lo - 1;
  if (bar >= st) {
    const int c1 = staged ? s_c1[bar - st] : a.ch_end[bo + bar] - skip;
    if (t < c1) {
      mk = 1;
      tg = (bar == nb - 1 && t == c1 - 1) ? (int64_t)a.eos : token(t + 1);   // EOS closes the last bar
    }
  }
  const int64_t o = (int64_t)b * T + t;
  a.inp[o] = token(t);
  a.tgt[o] = tg;
  a.mask[o] = mk;
  const uint8_t f = a.flags[tg];                         // :206-213, on ids instead of event strings
  a.chord_idx[o] = f & 1;
  a.melody_idx[o] = (f >> 1) & 1;
  if (t == 0) a.length[b] = seqlen < T ? seqlen : T;
}

// Stage-1 (lead-sheet) batches, reference stage1_compose/dataloader.py:408-445,469-520 (__getitem__ with the
// training configuration: first segment only, no augmentation) + collate_fn :194-255.  The store keeps, per piece,
// the sample's token sequence (events up to the last registered bar + the closing EOS / Bar token) and the number of
// positions E its first segment covers; input = tok[0:E], target = tok[1:E+1], truncated / PAD-filled to T; the
// chord / melody masks classify the TARGET event and are PAD-filled too (the reference pads them with pad_token).
struct Stage1Args {
  const int32_t* tokens;
  const int64_t* piece_off;   // [P + 1]
  const int32_t* seg_len;     // [P]: E
  const uint8_t* flags;
  const int32_t* sel_piece;   // [B]
  int64_t *inp, *tgt, *chord_idx, *melody_idx, *length;
  int T, pad;
};

__global__ void __launch_bounds__(256) stage1_batch_kernel(const Stage1Args a) {
  const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.T) return;
  const int p = a.sel_piece[b];
  const int64_t base = a.piece_off[p];
  const int E = a.seg_len[p];
  int64_t in = a.pad, tg = a.pad, ch = a.pad, me = a.pad;
  if (t < E) {
    in = a.tokens[base + t];
    tg = a.tokens[base + t + 1];
    const uint8_t f = a.flags[tg];
    ch = f & 1;
    me = (f >> 1) & 1;
  }
  const int64_t o = (int64_t)b * a.T + t;
  a.inp[o] = in;
  a.tgt[o] = tg;
  a.chord_idx[o] = ch;
  a.melody_idx[o] = me;
  if (t == 0) a.length[b] = E;                           // dec_seg_len: the UNtruncated segment length (:492)
}

}  // namespace

extern "C" int emo_stage1_batch(const int32_t* tokens, const int64_t* piece_off, const int32_t* seg_len,
                                const uint8_t* flags, const int32_t* sel_piece, int64_t* dec_inp, int64_t* dec_tgt,
                                int64_t* inp_chord, int64_t* inp_melody, int64_t* dec_seg_len, int B, int T,
                                int pad_token, void* stream) {
  EMO_REQUIRE(B >= 0 && T >= 1, "emo_stage1_batch: need T >= 1");
  if (B == 0) return EMO_OK;
  Stage1Args a;
  a.tokens = tokens; a.piece_off = piece_off; a.seg_len = seg_len; a.flags = flags; a.sel_piece = sel_piece;
  a.inp = dec_inp; a.tgt = dec_tgt; a.chord_idx = inp_chord; a.melody_idx = inp_melody; a.length = dec_seg_len;
  a.T = T; a.pad = pad_token;
  stage1_batch_kernel<<<dim3((T + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(a);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

extern "C" int emo_stage2_batch(const int32_t* tokens, const int64_t* piece_off, const int64_t* bar_off,
                                const int32_t* mel_start, const int32_t* ch_start, const int32_t* ch_end,
                                const uint8_t* flags, const int32_t* sel_piece, const int32_t* sel_stbar,
                                int64_t* dec_input, int64_t* dec_target, int64_t* track_mask, int64_t* chord_idx,
                                int64_t* melody_idx, int64_t* length, int B, int T, int pad_token, int eos_token,
                                int predict_key, void* stream) {
  EMO_REQUIRE(B >= 0 && T >= 2, "emo_stage2_batch: need T >= 2");
  if (B == 0) return EMO_OK;
  BatchArgs a;
  a.tokens = tokens; a.piece_off = piece_off; a.bar_off = bar_off; a.mel_start = mel_start; a.ch_start = ch_start;
  a.ch_end = ch_end; a.flags = flags; a.sel_piece = sel_piece; a.sel_stbar = sel_stbar;
  a.inp = dec_input; a.tgt = dec_target; a.mask = track_mask; a.chord_idx = chord_idx; a.melody_idx = melody_idx;
  a.length = length; a.T = T; a.pad = pad_token; a.eos = eos_token; a.predict_key = predict_key;
  stage2_batch_kernel<<<dim3((T + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(a);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
