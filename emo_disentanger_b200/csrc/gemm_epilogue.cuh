// Epilogue shared by the SIMT fp32 GEMM and the tcgen05 GEMM (include/emo_b200.h, emo_epilogue).
#pragma once
#include "common.cuh"

struct EpiParams {
  const float* bias;
  int act;
  const void* aux;
  void* aux_out;
  int64_t ld_aux;
  float aux_scale;
  uint32_t drop_thr;
  float keep_scale;
  uint64_t seed;
  const void* residual;
  int64_t ld_res;
  float alpha;
  int accumulate;
  const float* rowscale;
  float* colsum;     // optional fp32 [N]: += column sums of the stored values (fused bias gradient)
  const float* ln_gamma;   // A-operand LayerNorm prologue (decode rows): A := LN(A) * gamma + beta, also stored to ln_out
  const float* ln_beta;
  void* ln_out;
  int64_t ld_ln;
  int64_t n_total;   // logical N (dropout element index = m * n_total + n)
};

static inline EpiParams make_epi(const emo_epilogue* e, int64_t N) {
  EpiParams p;
  memset(&p, 0, sizeof(p));
  p.alpha = 1.f;
  p.keep_scale = 1.f;
  p.n_total = N;
  if (e) {
    p.bias = e->bias; p.act = e->act; p.aux = e->aux; p.aux_out = e->aux_out; p.ld_aux = e->ld_aux;
    p.aux_scale = e->aux_scale; p.drop_thr = emo_drop_thr(e->drop_p);
    p.keep_scale = 1.f / (1.f - e->drop_p); p.seed = e->seed; p.residual = e->residual; p.ld_res = e->ld_res;
    p.alpha = e->alpha; p.accumulate = e->accumulate; p.rowscale = (const float*)e->rowscale; p.colsum = e->colsum;
    p.ln_gamma = e->ln_gamma; p.ln_beta = e->ln_beta; p.ln_out = e->ln_out; p.ld_ln = e->ld_ln;
  }
  return p;
}

// everything except dropout / residual / store (those are done pairwise or vectorised by the caller)
template <typename TIn, typename TOut>
__device__ __forceinline__ float epi_pre(float acc, int64_t m, int64_t n, const EpiParams& p) {
  float v = acc * p.alpha;
  if (p.rowscale) v *= p.rowscale[m];
  if (p.bias) v += p.bias[n];
  switch (p.act) {
    case EMO_ACT_RELU: v = fmaxf(v, 0.f); break;
    case EMO_ACT_GELU_NEW:
      if (p.aux_out) reinterpret_cast<TOut*>(p.aux_out)[m * p.ld_aux + n] = from_f<TOut>(v);
      v = gelu_new_f(v);
      break;
    case EMO_ACT_RELU_MASK_BWD:
      v = (to_f(reinterpret_cast<const TIn*>(p.aux)[m * p.ld_aux + n]) != 0.f) ? v * p.aux_scale : 0.f;
      break;
    case EMO_ACT_GELU_NEW_BWD:
      v *= gelu_new_grad_f(to_f(reinterpret_cast<const TIn*>(p.aux)[m * p.ld_aux + n]));
      break;
    default: break;
  }
  return v;
}

template <typename TIn, typename TOut>
__device__ __forceinline__ float epi_full(float acc, int64_t m, int64_t n, const EpiParams& p) {
  float v = epi_pre<TIn, TOut>(acc, m, n, p);
  if (p.drop_thr) v = emo_drop_keep(p.seed, (uint64_t)(m * p.n_total + n), p.drop_thr) ? v * p.keep_scale : 0.f;
  if (p.residual) v += to_f(reinterpret_cast<const TOut*>(p.residual)[m * p.ld_res + n]);
  return v;
}
