// K4+K5: FAVOR+ random-feature map fused with the causal prefix-sum (linear attention).
//
// Replaces fast_transformers Favor.forward + CausalLinearAttention.forward + the native
// causal_dot_product extension (stage2_accompaniment/model/fast_transformer_decoder.py:28-38).
// One CTA per (batch, head) walks the sequence in chunks of C tokens; phi(q), phi(k) are recomputed
// per chunk in shared memory and never written to HBM.  Per chunk (V' = [v | 1 | 0..] is 80 wide so
// the normaliser rides along with the values):
//     U   = X . Om_s                 phi = [exp(U - o), exp(-U - o)],  o = |x|^2 s^2/2 + ln(128)/2
//     A   = tril(Phi_q Phi_k^T)                                  (intra-chunk causal part)
//     O'  = A V' + Phi_q S'          out = O'[:, :64] / (O'[:, 64] + 1e-6)
//     S' += Phi_k^T V'                                           (prefix state, fp32 registers)
// Backward is ONE reverse pass: it starts from the saved final state, un-does the state update
// chunk by chunk (S'_{c-1} = S'_c - Phi_k^T V'), and carries the reverse state R' = sum Phi_q^T G.
// All small products run through BlockGemm (tensor-core mma for bf16, fp32 FMA for the parity mode).
#include "favor_kernels.cuh"
using namespace FAVOR_NS;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
// Segment plan.  T is cut into segments of `sc` chunks; B*H*nseg CTAs run in parallel and are stitched together by
// the per-segment state sums (*_segsum kernels) turned into prefixes (favor_prefix_kernel).  Measured on B200
// (profiles/): every extra segment costs a fixed CTA prologue plus its share of the segsum pre-pass, so the best
// plan is the COARSEST one that still gives every SM a CTA: the backward (1 CTA / SM, 213 KB smem) gets
// nseg_b ~ SMs / (B*H) segments, the forward (2 CTAs / SM) twice as many while that still adds parallelism.
// Backward segment boundaries are a subset of the forward's (sc_b = ratio * sc_f), so the forward's prefix slots
// serve both.
int emo_favor_fwd2_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, void* out,
                          int64_t ld_out, float* den, const float* state_in, float* state_out, float* seg_states,
                          int nseg, int sc, int B, int T_, int H, cudaStream_t s);       // favor_fwd2.cu
int emo_favor_fwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, void* out,
                            int64_t ld_out, float* den, const float* state_in, float* state_out, float* seg_states,
                            int nseg, int sc, int B, int T_, int H, cudaStream_t s);     // favor_tc_fwd.cu (tcgen05 + TMA)
int emo_favor_bwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, const void* out,
                            const void* dout, int64_t ld_out, const float* den, const float* seg_states,
                            const float* seg_rstates, int nseg, int sc, int fwd_nseg, int ratio, void* dq, void* dk, void* dv,
                            int64_t ld_d, int B, int T_, int H, cudaStream_t s);     // favor_tc_bwd.cu (tcgen05 + TMA)
int emo_favor_bwd2_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, const void* out,
                          const void* dout, int64_t ld_out, const float* den, const float* seg_states,
                          const float* seg_rstates, int nseg, int sc, int fwd_nseg, int ratio, void* dq, void* dk, void* dv,
                          int64_t ld_d, int B, int T_, int H, cudaStream_t s);
#ifndef FAVOR_BWD2_DEFAULT
#define FAVOR_BWD2_DEFAULT 1
#endif
#ifndef FAVOR_FWD2_DEFAULT
#define FAVOR_FWD2_DEFAULT 1
#endif

// EMO_FAVOR_TC=0: A/B switch from the tcgen05 kernels (128-token chunks) back to the mma.sync ones
static int g_favor_tc = -1;
static int favor_tc_enabled() {
  if (g_favor_tc < 0) { const char* e = getenv("EMO_FAVOR_TC"); g_favor_tc = e ? atoi(e) : 1; }
  return g_favor_tc;
}
extern "C" void emo_favor_set_tc(int on) { g_favor_tc = on ? 1 : 0; }   // test / A-B hook
// EMO_FAVOR_TC_BWD=0: tcgen05 forward with the mma.sync backward (both read the same 128-token-chunk plan)
static int g_favor_tc_bwd = -1;
static int favor_tc_bwd_enabled() {
  if (g_favor_tc_bwd < 0) { const char* e = getenv("EMO_FAVOR_TC_BWD"); g_favor_tc_bwd = e ? atoi(e) : 1; }
  return g_favor_tc_bwd;
}
extern "C" void emo_favor_set_tc_bwd(int on) { g_favor_tc_bwd = on ? 1 : 0; }

struct FavorPlan { int nseg_f, sc_f, nseg_b, sc_b, ratio; };   // sc_* in chunks of FavorCfg<T>::C tokens
template <typename T> static FavorPlan favor_plan(int B, int T_, int H) {
  // the plan is made in units of the main kernels' chunk (128 tokens on the tcgen05 path) and handed out in units of
  // FavorCfg<T>::C, which is what the segment-sum pre-passes and the mma.sync kernels count in
  constexpr int C0 = FavorCfg<T>::C;
  const int C = (sizeof(T) == 2 && favor_tc_enabled()) ? 128 : C0;
  const int unit = C / C0;
  const int nchunk = (T_ + C - 1) / C;
  const int sms = emo_num_sms();
  const int64_t bh = (int64_t)B * H;
  int nb = (int)((sms + bh / 2) / bh);
  if (nb < 1) nb = 1;
  if (nb > nchunk) nb = nchunk;
  int sc_b = (nchunk + nb - 1) / nb;
  static int forced_f = -1, forced_b = -1;      // EMO_FAVOR_SEG_CHUNKS[_B]=<n>: tuning overrides
  if (forced_f < 0) { const char* e = getenv("EMO_FAVOR_SEG_CHUNKS"); forced_f = e ? atoi(e) : 0; }
  if (forced_b < 0) { const char* e = getenv("EMO_FAVOR_SEG_CHUNKS_B"); forced_b = e ? atoi(e) : 0; }
  FavorPlan p;
  const bool two = (bh * nb < 2 * (int64_t)sms) && sc_b >= 2;
  if (two && (sc_b & 1)) ++sc_b;
  p.sc_b = sc_b;
  p.sc_f = two ? sc_b / 2 : sc_b;
  if (forced_f > 0) {
    p.sc_f = forced_f < nchunk ? forced_f : nchunk;
    p.sc_b = p.sc_f;
    if (forced_b > 0 && forced_b % p.sc_f == 0) p.sc_b = forced_b;
  }
  p.ratio = p.sc_b / p.sc_f;
  p.nseg_f = (nchunk + p.sc_f - 1) / p.sc_f;
  p.nseg_b = (nchunk + p.sc_b - 1) / p.sc_b;
  p.sc_f *= unit;
  p.sc_b *= unit;
  return p;
}

// number of [128,80] fp32 slots per (b,h) of the seg_states / seg_rstates workspaces: the forward's segment
// prefixes + the total
extern "C" int emo_favor_nseg(int B, int T, int H, int dtype) {
  if (B * H <= 0 || T <= 0) return 2;
  FavorPlan p = dtype == EMO_BF16 ? favor_plan<bf16>(B, T, H) : favor_plan<float>(B, T, H);
  return p.nseg_f + 1;
}

template <typename T>
static int favor_fwd_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, void* out,
                            int64_t ld_out, float* den, const float* state_in, float* state_out, float* seg_states,
                            int B, int T_, int H, cudaStream_t s) {
  constexpr int C = FavorCfg<T>::C;
  size_t smem = sizeof(FavorSmemFwd<T, C>);
  int nseg = 1, sc = (T_ + C - 1) / C;
  if (seg_states) { FavorPlan pl = favor_plan<T>(B, T_, H); nseg = pl.nseg_f; sc = pl.sc_f; }
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_segsum_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FavorSmemSeg<T, C>)));
    configured = true;
  }
  if (seg_states && nseg > 1) {
    favor_segsum_kernel<T><<<B * H * (nseg - 1), BG_THREADS, sizeof(FavorSmemSeg<T, C>), s>>>((const T*)k, (const T*)v, ld, omega, seg_states, nseg, sc, T_, H);
    EMO_LAUNCH_CHECK();
    if (nseg > 2) {          // with two segments slot 1 already is the prefix of segment 1
      const int64_t n = (int64_t)B * H * FM * FV / 4;
      favor_prefix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seg_states, nseg, 0, (int64_t)B * H);
      EMO_LAUNCH_CHECK();
    }
  }
  static int fwd2 = -1;       // EMO_FAVOR_FWD2=0: A/B switch back to the block-GEMM forward
  if (fwd2 < 0) { const char* e = getenv("EMO_FAVOR_FWD2"); fwd2 = e ? atoi(e) : FAVOR_FWD2_DEFAULT; }
  if (sizeof(T) == 2 && favor_tc_enabled())
    return emo_favor_fwd_tc_launch(q, k, v, ld, omega, out, ld_out, den, state_in, state_out, seg_states, nseg, (sc + 1) / 2, B, T_, H, s);
  if (sizeof(T) == 2 && fwd2)
    return emo_favor_fwd2_launch(q, k, v, ld, omega, out, ld_out, den, state_in, state_out, seg_states, nseg, sc, B, T_, H, s);
  favor_fwd_kernel<T><<<B * H * nseg, BG_THREADS, smem, s>>>((const T*)q, (const T*)k, (const T*)v, ld, omega, (T*)out, ld_out,
                                                             den, state_in, state_out, seg_states, nseg, sc, T_, H);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
template <typename T>
static int favor_bwd_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega,
                            const void* out, const void* dout, int64_t ld_out, const float* den,
                            const float* seg_states, float* seg_rstates, void* dq, void* dk, void* dv, int64_t ld_d,
                            int B, int T_, int H, cudaStream_t s) {
  constexpr int C = FavorCfg<T>::C;
  size_t smem = sizeof(FavorSmemBwd<T, C>);
  const FavorPlan pl = favor_plan<T>(B, T_, H);
  const int nseg = pl.nseg_b, sc = pl.sc_b;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_bwd_segsum_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FavorSmemSeg<T, C>)));
    configured = true;
  }
  if (nseg > 1) {
    favor_bwd_segsum_kernel<T><<<B * H * nseg, BG_THREADS, sizeof(FavorSmemSeg<T, C>), s>>>((const T*)q, ld, omega, (const T*)out, (const T*)dout,
                                                                      ld_out, den, seg_rstates, nseg, sc, T_, H);
    EMO_LAUNCH_CHECK();
    const int64_t n = (int64_t)B * H * FM * FV / 4;
    favor_prefix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seg_rstates, nseg, 1, (int64_t)B * H);
    EMO_LAUNCH_CHECK();
  }
  static int bwd2 = -1;       // EMO_FAVOR_BWD2=0/1: A/B switch between the block-GEMM and the register-resident backward
  if (bwd2 < 0) { const char* e = getenv("EMO_FAVOR_BWD2"); bwd2 = e ? atoi(e) : FAVOR_BWD2_DEFAULT; }
  if (sizeof(T) == 2 && favor_tc_enabled() && favor_tc_bwd_enabled())
    return emo_favor_bwd_tc_launch(q, k, v, ld, omega, out, dout, ld_out, den, seg_states, seg_rstates, nseg, sc / 2, pl.nseg_f, pl.ratio,
                                   dq, dk, dv, ld_d, B, T_, H, s);
  if (sizeof(T) == 2 && bwd2)
    return emo_favor_bwd2_launch(q, k, v, ld, omega, out, dout, ld_out, den, seg_states, seg_rstates, nseg, sc, pl.nseg_f, pl.ratio,
                                 dq, dk, dv, ld_d, B, T_, H, s);
  favor_bwd_kernel<T><<<B * H * nseg, BG_THREADS, smem, s>>>((const T*)q, (const T*)k, (const T*)v, ld, omega, (const T*)out,
                                                             (const T*)dout, ld_out, den, seg_states, seg_rstates, nseg, sc,
                                                             pl.nseg_f, pl.ratio, (T*)dq, (T*)dk, (T*)dv, ld_d, T_, H);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int emo_favor_fwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const float* omega,
                             void* out, int64_t ld_out, float* den, const float* state_in, float* state_out,
                             float* seg_states, int B, int T, int H, int dtype, void* stream) {
  int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && aligned16(omega), "emo_favor_fwd: pointers must be 16-byte aligned");
  EMO_REQUIRE((ld_qkv * esz) % 16 == 0 && (ld_out * esz) % 16 == 0, "emo_favor_fwd: row strides must be 16-byte multiples");
  if (B * H == 0 || T == 0) return EMO_OK;
  if (dtype == EMO_BF16) return favor_fwd_launch<bf16>(q, k, v, ld_qkv, omega, out, ld_out, den, state_in, state_out, seg_states, B, T, H, (cudaStream_t)stream);
  return favor_fwd_launch<float>(q, k, v, ld_qkv, omega, out, ld_out, den, state_in, state_out, seg_states, B, T, H, (cudaStream_t)stream);
}

extern "C" int emo_favor_bwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const float* omega,
                             const void* out, const void* dout, int64_t ld_out, const float* den,
                             const float* seg_states, float* seg_rstates, void* dq, void* dk, void* dv,
                             int64_t ld_dqkv, int B, int T, int H, int dtype, void* stream) {
  int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && aligned16(dout) && aligned16(dq) &&
                  aligned16(dk) && aligned16(dv) && aligned16(omega), "emo_favor_bwd: pointers must be 16-byte aligned");
  EMO_REQUIRE((ld_qkv * esz) % 16 == 0 && (ld_out * esz) % 16 == 0 && (ld_dqkv * esz) % 16 == 0,
              "emo_favor_bwd: row strides must be 16-byte multiples");
  EMO_REQUIRE(den != nullptr && seg_states != nullptr && seg_rstates != nullptr, "emo_favor_bwd: den, seg_states and seg_rstates are required");
  if (B * H == 0 || T == 0) return EMO_OK;
  if (dtype == EMO_BF16)
    return favor_bwd_launch<bf16>(q, k, v, ld_qkv, omega, out, dout, ld_out, den, seg_states, seg_rstates, dq, dk, dv, ld_dqkv, B, T, H, (cudaStream_t)stream);
  return favor_bwd_launch<float>(q, k, v, ld_qkv, omega, out, dout, ld_out, den, seg_states, seg_rstates, dq, dk, dv, ld_dqkv, B, T, H, (cudaStream_t)stream);
}

extern "C" int emo_favor_step(const void* q, const void* k, const void* v, int64_t ld_qkv, const float* omega,
                              float* state, void* out, int64_t ld_out, int B, int H, int dtype, void* stream) {
  if (B * H == 0) return EMO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == EMO_BF16)
    EMO_CHECK_CUDA(emo_launch_dep(favor_step_kernel<bf16>, dim3(B * H), dim3(256), 0, s, (const bf16*)q, (const bf16*)k, (const bf16*)v,
                                  ld_qkv, omega, state, (bf16*)out, ld_out, H));
  else
    EMO_CHECK_CUDA(emo_launch_dep(favor_step_kernel<float>, dim3(B * H), dim3(256), 0, s, (const float*)q, (const float*)k,
                                  (const float*)v, ld_qkv, omega, state, (float*)out, ld_out, H));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
