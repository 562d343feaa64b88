/* libemo_b200.so -- C ABI of the B200-native EMO-Disentanger hot path.
 *
 * The reference (Yuer867/EMO-Disentanger) has no FFI of its own: its operator boundary is the
 * Python nn.Module surface plus the third-party ops it calls.  Each entry point below replaces
 * one such operator; the citation gives the reference call site (paths relative to the
 * reference repo) or the third-party operator it stands in for.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory owned by the caller
 *     (PyTorch tensors used as storage), including workspaces; the library keeps no tensor memory.
 *   - stream-ordered and asynchronous: no internal synchronisation; `stream` is a cudaStream_t
 *     passed as void*.
 *   - returns 0 on success, non-zero emo_status otherwise; message via emo_last_error()
 *     (thread-local).  Never throws, never exits.
 *   - dtype: EMO_F32 or EMO_BF16 for activations; parameters / optimizer state are fp32 masters
 *     (bf16 shadow copies are produced by emo_adam_step / emo_cast).
 *   - model geometry is the reference's: d_model 512, 8 heads x 64, FAVOR n_dims 128.
 */
#ifndef EMO_B200_H
#define EMO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { EMO_OK = 0, EMO_ERR_ARG = 1, EMO_ERR_CUDA = 2, EMO_ERR_UNSUPPORTED = 3 } emo_status;
typedef enum { EMO_F32 = 0, EMO_BF16 = 1 } emo_dtype;

int emo_version(void);
const char* emo_last_error(void);

/* ---- A1: embedding front-end --------------------------------------------------------------
 * stage2_accompaniment/model/music_performer.py:51-62 + transformer_helpers.py:57-63,81-87
 * (identical in music_gpt2.py:71-82); stage1_compose/model/plain_transformer.py:61-62.
 * out[b,t,:] = dropout( (E_tok[tok] + E_seg[seg]) * scale + pe[t] ).  tok/seg are int64 with
 * element (b,t) at tok[b*stride_b + t*stride_t] (stage 1 passes its [T,B] layout by strides).
 * e_seg / seg / pe may be NULL.  out is [B,T,d] contiguous. */
int emo_embed_fwd(const int64_t* tok, const int64_t* seg, int64_t stride_b, int64_t stride_t,
                  const float* e_tok, const float* e_seg, const float* pe, void* out,
                  int B, int T, int d, float scale, float drop_p, uint64_t seed,
                  int out_dtype, void* stream);
/* decode step of the same front-end (stage2_accompaniment/inference.py:252-272 feeds the model one new
 * token per iteration): one row per sequence, out[b,:] = (E_tok[tok[b]] + E_seg[seg[b]]) * scale +
 * pe[pos[b]], with tok / seg / pos read from DEVICE memory (ragged batch, CUDA-graph capturable). */
int emo_embed_rows(const int64_t* tok, const int64_t* seg, const int64_t* pos, const float* e_tok,
                   const float* e_seg, const float* pe, void* out, int rows, int d, float scale,
                   int64_t* pos_advance /* NULL, or where pos[row] + 1 is written (may be pos) */,
                   int out_dtype, void* stream);
/* ---- 8f rank 1: stage-2 training batches from a GPU-resident token store ---------------------------------
 * stage2_accompaniment/dataloader.py:178-231 (REMISkylineToMidiTransformerDataset.__getitem__) + DataLoader
 * collation + the H2D copies of train.py:44-50.  tokens: every piece's event ids, concatenated (int32);
 * piece_off [P+1] / bar_off [P+1] (int64) index tokens / the per-bar tables; mel_start[b] = melody_pos[b][0],
 * [ch_start[b], ch_end[b]) = chord_pos[b] (the Full-track span of bar b); flags [V]: bit 0 = 'Chord_*', bit 1 =
 * 'Note_*'.  sel_piece / sel_stbar [B]: the sample picks (piece index, start bar).  Outputs [B, T] int64:
 * dec_input (header + events from the start bar, PAD-filled / truncated to T), dec_target (next token inside the
 * Full-track spans, EOS closing the last bar, PAD elsewhere), track_mask, chord_idx, melody_idx; length [B].
 * Integer work: bit-exact against the reference. */
int emo_stage2_batch(const int32_t* tokens, const int64_t* piece_off, const int64_t* bar_off,
                     const int32_t* mel_start, const int32_t* ch_start, const int32_t* ch_end,
                     const uint8_t* flags, const int32_t* sel_piece, const int32_t* sel_stbar,
                     int64_t* dec_input, int64_t* dec_target, int64_t* track_mask, int64_t* chord_idx,
                     int64_t* melody_idx, int64_t* length, int B, int T, int pad_token, int eos_token,
                     int predict_key, void* stream);

/* ---- stage-1 dataset: SkylineFullSongTransformerDataset.__getitem__ + collate_fn -------------------------
 * stage1_compose/dataloader.py:408-445,469-520,194-255 as train.py:230-262 configures it (first segment of a piece,
 * no augmentation).  tokens / piece_off: per piece the sample's token sequence (events up to the last registered bar
 * + the closing EOS / Bar id); seg_len[p] = positions covered by the first segment.  Writes [B, T] int64 dec_inp,
 * dec_tgt (next token), inp_chord / inp_melody (the TARGET is a Chord_* / Note_* event), all PAD-filled past the
 * segment, and dec_seg_len [B] (untruncated).  The unused encoder features (enc_*) are not produced. */
int emo_stage1_batch(const int32_t* tokens, const int64_t* piece_off, const int32_t* seg_len,
                     const uint8_t* flags, const int32_t* sel_piece, int64_t* dec_inp, int64_t* dec_tgt,
                     int64_t* inp_chord, int64_t* inp_melody, int64_t* dec_seg_len, int B, int T,
                     int pad_token, void* stream);

/* ---- A11: the whole per-token step of the stage-2 Performer in ONE kernel --------------------------------
 * stage2_accompaniment/inference.py:252-272 (one model call per generated token).  One thread-block cluster (16
 * CTAs) per sequence: embedding row -> 12 post-LN layers (qkv GEMV, FAVOR+ recurrent step, out-proj + residual, LN,
 * FFN1 + ReLU, FFN2 + residual, LN) -> logits; the 61 phase boundaries are hardware cluster barriers and the
 * weights of a warp's next phase are in flight before it enters the barrier.  bf16 weights / activations, fp32
 * accumulation; bit-identical to the same step issued as emo_embed_rows / emo_gemm / emo_favor_step calls.
 * w_bf16 / w_f32: the model's flat parameter buffer in bf16 and fp32 (same element offsets); layer_offs [n_layer][12]
 * int64 element offsets (DEVICE memory): Wqkv Wo W1 W2 | bqkv bo b1 b2 | norm1.w norm1.b norm2.w norm2.b;
 * off_seg < 0 / pe == NULL: no segment embedding / positional encoding.  tok, seg, pos: int64 [batch] on the
 * device (pos is advanced by one).  state [n_layer, batch, 8, 128, 80] fp32 (updated in place), omegas
 * [n_layer, 64, 64] fp32, scratch: batch * 6144 bf16, logits [batch, ld_logits] fp32. */
int emo_performer_decode_step(const void* w_bf16, const float* w_f32, const int64_t* layer_offs,
                              int64_t off_tok, int64_t off_seg, int64_t off_outw, int64_t off_outb,
                              const float* pe, const float* omegas, float* state, const int64_t* tok,
                              const int64_t* seg, int64_t* pos, void* scratch, float* logits, int n_layer,
                              int batch, int n_token, int ld_logits, float emb_scale, void* stream);
/* 1: the decode-step kernels (emo_embed_rows, the M <= 8 path of emo_gemm, emo_favor_step) are launched as
 * programmatic dependent launches -- each may start while its predecessor on the stream drains (its weight
 * prefetch overlaps the predecessor's tail) and waits (griddepcontrol.wait) before touching its inputs.  Meant
 * to be switched on around the capture of the per-token CUDA graph; 0 (default) = ordinary launches. */
void emo_set_pdl(int on);
/* d_e_tok[tok] += dout*scale*mask (fp32 atomics); rows == pad_idx get no gradient (pass -1 for
 * none: nn.Embedding(padding_idx) in stage1 transformer_helpers.py:104-108). */
int emo_embed_bwd(const int64_t* tok, const int64_t* seg, int64_t stride_b, int64_t stride_t,
                  const void* dout, float* d_e_tok, float* d_e_seg, int B, int T, int d,
                  float scale, float drop_p, uint64_t seed, int64_t pad_idx, int dtype,
                  void* stream);

/* ---- A3/A7/A9: LayerNorm (eps 1e-5) over the last dim d=512 -------------------------------
 * fast_transformers TransformerEncoderLayer.norm1/norm2, HF GPT2Block.ln_1/ln_2,
 * optimus_txl_decoder.py:52,318.  Saves mean/rstd (fp32 [rows]) for backward. */
int emo_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
               float* rstd, int64_t rows, int d, float eps, int dtype, void* stream);
/* y = LN(x + res) with the sum formed in fp32 inside the kernel: x is a projection's output WITHOUT its residual
 * (out_projection / linear2 after dropout), res the sub-layer's input (TransformerEncoderLayer.forward:
 * `x = norm1(x + dropout(attn))`, `norm2(x + y)`).  sum_out (may be NULL) receives x + res rounded to the
 * compute dtype: the tensor emo_ln_bwd re-normalises. */
int emo_ln_res_fwd(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta,
                   void* y, float* mean, float* rstd, int64_t rows, int d, float eps, int dtype,
                   void* stream);
/* dx = LNgrad(dy) (+ add_in if non-NULL).  If dx_drop != NULL also writes
 * dx_drop = dx * dropmask(seed)/(1-p) (gradient entering a `x + dropout(proj)` branch).
 * dgamma/dbeta (fp32 [d]) are ACCUMULATED (atomics).  dxsum (fp32 [d], may be NULL) += column sums of
 * dx_drop (of dx when dx_drop is NULL): the bias gradient of the projection this gradient flows into
 * (out_projection / linear2 / c_proj), so no separate pass over the tensor is needed. */
int emo_ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
               const float* gamma, const void* add_in, void* dx, void* dx_drop, float drop_p,
               uint64_t seed, float* dgamma, float* dbeta, float* dxsum, int64_t rows, int d, int dtype,
               void* stream);
/* y = dropmask(seed)/(1-p) * x  (elementwise helper for branches without an LN in between) */
int emo_dropout_apply(const void* x, void* y, int64_t n, float drop_p, uint64_t seed, int dtype,
                      void* stream);

/* ---- A3/A4/A7/A8/A9: dense projections -----------------------------------------------------
 * nn.Linear / HF Conv1D call sites: AttentionLayer.{query,key,value,out}_projection,
 * TransformerEncoderLayer.linear1/2, GPT2 c_attn/c_proj/c_fc, qkv_net/r_net/o_net/CoreNet,
 * dec_out_proj (music_performer.py:65).  bf16 inputs run on tcgen05 tensor cores
 * (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue); fp32 inputs run a SIMT fp32 kernel (the
 * 1e-3 parity mode).  Three contractions, row-major operands, leading dims in elements:
 *   EMO_GEMM_NT: C[M,N] = A[M,K] . B[N,K]^T     (forward with nn.Linear weight [out,in])
 *   EMO_GEMM_NN: C[M,N] = A[M,K] . B[K,N]       (dgrad; forward with Conv1D weight [in,out])
 *   EMO_GEMM_TN: C[M,N] = A[K,M]^T . B[K,N]     (wgrad: reduction over tokens)
 * Epilogue, in this order: v = alpha*acc + bias[n]; activation; dropout(seed); + residual;
 * then either store (out_dtype) or atomically accumulate into fp32 C (accumulate=1). */
typedef enum { EMO_GEMM_NT = 0, EMO_GEMM_NN = 1, EMO_GEMM_TN = 2 } emo_gemm_op;
typedef enum {
  EMO_ACT_NONE = 0,
  EMO_ACT_RELU = 1,
  EMO_ACT_GELU_NEW = 2,      /* writes pre-activation to aux_out if non-NULL               */
  EMO_ACT_RELU_MASK_BWD = 3, /* v *= (aux[m,n] != 0) ? aux_scale : 0   (relu+dropout bwd)    */
  EMO_ACT_GELU_NEW_BWD = 4   /* v *= gelu_new'(aux[m,n])                                     */
} emo_act;
typedef struct {
  const float* bias;    /* [N] fp32 or NULL                                                   */
  int act;              /* emo_act                                                            */
  const void* aux;      /* [M, ld_aux] in the input dtype (acts 3, 4)                         */
  void* aux_out;        /* [M, ld_aux] in the output dtype (act 2), may be NULL               */
  int64_t ld_aux;
  float aux_scale;      /* act 3: 1/(1-p)                                                     */
  float drop_p;         /* forward dropout after the activation (0 = off)                     */
  uint64_t seed;
  const void* residual; /* [M, ld_res] in the output dtype, added last; may be NULL           */
  int64_t ld_res;
  float alpha;
  int accumulate;       /* 1: C (fp32) += result                                             */
  const void* rowscale; /* optional fp32 [M]: v *= rowscale[m] before bias (unused = NULL)    */
  float* colsum;        /* optional fp32 [N]: colsum[n] += sum_m C[m,n] of the values stored (the
                           bias gradient of the producing layer, fused into its dgrad GEMM);
                           bf16 tensor-core path with N % 64 == 0 only, else EMO_ERR_UNSUPPORTED */
  const float* ln_gamma; /* optional A-operand prologue (bf16, K == 512): A := LayerNorm(A)*gamma+beta   */
  const float* ln_beta;  /* (eps 1e-5) before the product; the normalised rows are also written to       */
  void* ln_out;          /* ln_out [M, ld_ln] (input dtype; required).  Fused into the decode-rows GEMV   */
  int64_t ld_ln;         /* (M <= 8); otherwise emo_gemm runs emo_ln_fwd into ln_out first.               */
} emo_epilogue;
int emo_gemm(int op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
             int64_t ldb, void* C, int64_t ldc, int in_dtype, int out_dtype,
             const emo_epilogue* epi, void* stream);
/* column sums: out[n] += sum_m x[m,n]  (bias gradients), fp32 accumulate with atomics */
int emo_colsum(const void* x, int64_t ld, int64_t M, int64_t N, float* out, int dtype,
               void* stream);

/* ---- A5/A6: FAVOR+ feature map + causal linear attention ----------------------------------
 * fast_transformers.feature_maps.Favor.forward + attention.CausalLinearAttention.forward +
 * causal_product.causal_dot_product (the only native extension on the reference path), as
 * built by stage2_accompaniment/model/fast_transformer_decoder.py:28-38.
 * q,k,v: [B,T,H,64] with token row stride ld_qkv (slices of one packed [B*T, 3*512] buffer);
 * omega [64,64] fp32 (row = input dim, col = feature); phi is recomputed in-kernel and never
 * written to HBM.  out [B,T,H*64] (ld_out); den [B,T,H] fp32 = phi(q).cumsum(phi(k)) + 1e-6;
 * state_in (NULL = start of sequence) / state_out (may be NULL; may alias state_in) [B,H,128,80] fp32:
 * prefix state [sum phi(k) v^T | sum phi(k) | 0] before / after these T tokens (decode appends blocks
 * of tokens -- a lead-sheet bar -- to a running state, stage2_accompaniment/inference.py:293-307).
 * seg_states (NULL = one sequential pass per (b,h)): workspace [B,H,S,128,80] fp32, S = emo_favor_nseg(B,T,H,dtype)
 * slots (S-1 segments + the total).  When given, the sequence is cut into S-1 segments that run in parallel
 * (B*H*(S-1) CTAs): a first kernel writes every segment's local state sum there, a tiny scan turns them into
 * exclusive prefixes in place, the main kernel starts each segment from its prefix.  The same buffer is what
 * emo_favor_bwd needs. */
int emo_favor_nseg(int B, int T, int H, int dtype);
int emo_favor_fwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const float* omega,
                  void* out, int64_t ld_out, float* den, const float* state_in, float* state_out,
                  float* seg_states, int B, int T, int H, int dtype, void* stream);
/* reverse-scan backward, segment-parallel like the forward: seg_states = the forward's workspace (required,
 * from a forward call with state_in == NULL); seg_rstates = scratch of the same shape for the reverse state
 * sum phi(q)^T G of every segment. */
int emo_favor_bwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const float* omega,
                  const void* out, const void* dout, int64_t ld_out, const float* den,
                  const float* seg_states, float* seg_rstates, void* dq, void* dk, void* dv, int64_t ld_dqkv,
                  int B, int T, int H, int dtype, void* stream);
/* one decode step per sequence (recurrent form): state [B,H,128,80] fp32 updated in place;
 * q,k,v rows for the new token [B,H,64] (row stride ld_qkv per sequence). */
int emo_favor_step(const void* q, const void* k, const void* v, int64_t ld_qkv, const float* omega,
                   float* state, void* out, int64_t ld_out, int B, int H, int dtype, void* stream);

/* ---- A7: GPT-2 causal softmax attention ---------------------------------------------------
 * HF GPT2Attention._attn (4.28.0): softmax(q k^T / 8 + causal) v with attention-prob dropout.
 * q [B,Tq,H,64], k,v [B,Tk,H,64] (row strides ld_q / ld_kv); query i sees keys j <= i + (Tk-Tq).
 * lse [B,H,Tq] fp32 saved for backward. */
int emo_attn_fwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv,
                 void* out, int64_t ld_out, float* lse, int B, int Tq, int Tk, int H, float scale,
                 float drop_p, uint64_t seed, int dtype, void* stream);
int emo_attn_bwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv,
                 const void* out, const void* dout, int64_t ld_out, const float* lse, void* dq,
                 void* dk, void* dv, int64_t ld_dq, int64_t ld_dkv, int B, int Tq, int Tk, int H,
                 float scale, float drop_p, uint64_t seed, int dtype, void* stream);

/* Decode step of the GPT-2 attention for a ragged batch (one new token per sequence): appends this token's k | v row
 * (qkv row b = [q | k | v], 3*H*64 values) to sequence b's cache [max_len][2*H*64] at pos[b] (int64, device) and attends
 * over keys 0..pos[b].  Replaces the full-prefix re-run of HF GPT2Attention per sampled token
 * (stage2_accompaniment/inference.py:252-272).  Static shapes: capturable in a CUDA graph. */
int emo_attn_decode_step(const void* qkv, int64_t ld_qkv, void* kv_cache, int64_t max_len, const int64_t* pos,
                         void* out, int64_t ld_out, int B, int H, float scale, int dtype, void* stream);

/* ---- A9: stage-1 relative-position attention ----------------------------------------------
 * optimus_txl_decoder.py:305-387: score(i,j) = ((q_i+r_w_bias).k_j + (q_i+r_r_bias).r_{dist})/8,
 * dist = i + mlen - j (>= 0 visible), softmax, dropatt(drop_p, seed), renormalise /(sum+1e-8), . v
 * (drop + renormalise = softmax over the randomly kept keys; a row whose keys were all dropped gives 0).
 * dr [Tk,H,64], d_r_w_bias, d_r_r_bias [H,64] are fp32 and ACCUMULATED (atomics).
 * q [B,Tq,H,64]; k,v [B,Tk,H,64]; r [Tk,H,64] with row p holding distance Tk-1-p (as r_net of
 * pos_emb for pos_seq = Tk-1..0, :792-796); biases [H,64] fp32. */
/* Decode step of the stage-1 attention (one new token per sequence): K / V of past positions are cached instead of being
 * re-derived from the hidden-state memory on every step (stage1_compose/inference_utils.py:100-104 -> optimus_txl_decoder
 * .py:702-722, 336-367); the new token attends over the last mem_len + 1 positions.  rtab [mem_len + 1][H*64] (compute
 * dtype) = r_net(pos_emb(distance)) of this layer, row = distance; cache [B][cap][2*H*64]; pos int64 [B] on the device. */
int emo_relattn_decode_step(const void* qkv, int64_t ld_qkv, void* kv_cache, int64_t cap, const int64_t* pos,
                            const void* rtab, const float* r_w_bias, const float* r_r_bias, int mem_len, void* out,
                            int64_t ld_out, int B, int H, float scale, int dtype, void* stream);
int emo_relattn_fwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv,
                    const void* r, int64_t ld_r, const float* r_w_bias, const float* r_r_bias,
                    void* out, int64_t ld_out, float* lse, int B, int Tq, int Tk, int H,
                    float scale, float drop_p, uint64_t seed, int dtype, void* stream);
int emo_relattn_bwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv,
                    const void* r, int64_t ld_r, const float* r_w_bias, const float* r_r_bias,
                    const void* out, const void* dout, int64_t ld_out, const float* lse,
                    void* dq, void* dk, void* dv, int64_t ld_dq, int64_t ld_dkv, float* dr,
                    float* d_r_w_bias, float* d_r_r_bias, int B, int Tq, int Tk, int H,
                    float scale, float drop_p, uint64_t seed, int dtype, void* stream);

/* ---- A8: cross-entropy over the vocabulary -------------------------------------------------
 * compute_loss (music_performer.py:72-81, music_gpt2.py:94-103, plain_transformer.py:82-93):
 * mean NLL over targets != ignore_index.  logits fp32 [rows, ld]; tgt int64 with element r at
 * tgt[(r / tgt_inner) * tgt_stride_outer + (r % tgt_inner) * tgt_stride_inner].
 * emo_ce_count: count[0] += #(tgt != ignore).   emo_ce_fwd_bwd: loss_sum[0] += sum NLL;
 * ncorrect[0] += #(argmax == tgt) over non-ignored rows; pred[r] = argmax (may be NULL);
 * dlogits (may be NULL; dtype dl_dtype, ld_dl) = gscale * (softmax - onehot) / count[0], zero on
 * ignored rows and on pad columns [V, ld_dl). */
int emo_ce_count(const int64_t* tgt, int64_t rows, int64_t tgt_inner, int64_t tgt_stride_outer,
                 int64_t tgt_stride_inner, int64_t ignore_index, float* count, void* stream);
int emo_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* tgt, int64_t rows,
                   int64_t tgt_inner, int64_t tgt_stride_outer, int64_t tgt_stride_inner, int V,
                   int64_t ignore_index, const float* count, float gscale, float* loss_sum,
                   float* ncorrect, int32_t* pred, void* dlogits, int64_t ld_dl, int dl_dtype,
                   void* stream);

/* ---- A10: clip_grad_norm_ + Adam on flat fp32 buffers --------------------------------------
 * stage2_accompaniment/train.py:79-81,318-322; stage1_compose/train.py:63-65,287-293.
 * emo_sumsq: out[0] += sum g^2.  emo_adam_step: coef = min(1, max_norm/(sqrt(gnorm_sq[0])+1e-6))
 * (max_norm <= 0 disables clipping), torch.optim.Adam update (no weight decay, no amsgrad) with
 * bias corrections for `step` (1-based); optionally emits the bf16 shadow copy of the params and
 * zeroes the gradient buffer for the next step. */
int emo_sumsq(const float* g, int64_t n, float* out, void* stream);
int emo_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                  float beta1, float beta2, float eps, int64_t step, const float* gnorm_sq,
                  float max_norm, float grad_scale, int zero_grad, void* stream);
int emo_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream);

/* ---- A12: temperature + nucleus sampling ---------------------------------------------------
 * stage2_accompaniment/inference.py:71-100, stage1_compose/inference_utils.py:14-41.
 * logits fp32 [rows, ld] (V <= 1024).  greedy != 0 -> argmax (bit-exact decode mode; lowest
 * index wins ties like numpy.argmax).  Otherwise softmax(l/t), descending sort, cut at the second
 * index whose cumulative mass exceeds top_p (top-3 fallback), renormalise, inverse-CDF draw with
 * the caller-provided uniform u[row] in [0,1).  out int64 [rows]; status[row] = 1 when exactly
 * one index exceeded top_p (the reference raises IndexError there).
 * banned (NULL = off; uint8 [rows, V], non-zero = inadmissible): grammar-constrained draw.  The
 * decode loops reject inadmissible samples and re-draw (inference.py:279-310, inference_utils.py:80-118), i.e. they
 * sample from the nucleus candidates restricted to the admissible tokens; with `banned` those candidates get zero
 * mass inside the kernel -- the same distribution in one draw, no wasted model calls.  status 2 = every candidate is
 * inadmissible. */
int emo_sample(const float* logits, int64_t ld, int rows, int V, float temperature, float top_p,
               const float* u, int greedy, int64_t* out, int32_t* status, const uint8_t* banned, void* stream);
/* the same with one temperature per row (fp32 [rows] on the device): a batch of sequences decoded in lockstep with
 * different settings (the four emotion quadrants of stage2_accompaniment/inference.py:455-462) */
int emo_sample_rows(const float* logits, int64_t ld, int rows, int V, const float* temperature_rows, float top_p,
                    const float* u, int greedy, int64_t* out, int32_t* status, const uint8_t* banned,
                    void* stream);

/* ---- A3 fused: projection + dropout + residual + LayerNorm of the post-LN encoder layer -----------------------
 * fast_transformers TransformerEncoderLayer.forward (`x = norm1(x + dropout(attn_out))`, `norm2(x + dropout(ff))`),
 * N = d_model = 512, bf16 operands:   s = res + dropout(A[M,K] . W[512,K]^T + bias);  y = LN(s) * gamma + beta.
 * s is formed and normalised in fp32 inside the kernel (a CTA pair owns whole rows: 512 TMEM columns); mean / rstd
 * (fp32 [M], may be NULL) and sum_out (bf16 [M, ld_sum], may be NULL: the tensor emo_ln_bwd re-normalises) are kept for
 * the backward.  Same dropout mask as emo_gemm with the same seed (element index m * 512 + n).  Equivalent to emo_gemm
 * (NT, bias, drop_p) followed by emo_ln_res_fwd, without the 620 MB that pair moves through HBM per call. */
int emo_gemm_ln_res(int64_t M, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                    float drop_p, uint64_t seed, const void* res, int64_t ld_res, const float* gamma,
                    const float* beta, float eps, void* y, int64_t ldy, void* sum_out, int64_t ld_sum, float* mean,
                    float* rstd, void* stream);

/* ---- A11/A12 fused: logits projection + sampler of the decode loop ---------------------------
 * stage2_accompaniment/inference.py:272-276 (`logits = model(...)[-1]` -> `temperature` -> `nucleus`),
 * stage1_compose/inference_utils.py:66-72.  One launch per generated token: a thread-block cluster per sequence
 * computes logits[row, :V] = LN?(x[row, :K]) . W[V, K]^T + bias (bf16 operands, fp32 out; K == 512), stores them to
 * `logits` (fp32 [rows, ld]: the host-side reject-and-redraw loops read them) and draws the token from an on-chip
 * copy exactly as emo_sample / emo_sample_rows would (same arguments, same status words).  ln_gamma / ln_beta (NULL =
 * off): LayerNorm of the hidden row first (the post-LN Performer's last norm2, fast_transformer_decoder.py:62-67).
 * The logits are bit-identical to emo_gemm (NT, rows <= 8) on the same operands, hence so are the greedy tokens.
 * temperature_rows (NULL = use `temperature`): one temperature per row. */
int emo_logits_sample(const void* x, int64_t ldx, const float* ln_gamma, const float* ln_beta, const void* W,
                      int64_t ldw, const float* bias, int rows, int V, int K, float* logits, int64_t ld,
                      float temperature, const float* temperature_rows, float top_p, const float* u, int greedy,
                      int64_t* out, int32_t* status, const uint8_t* banned, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMO_B200_H */
