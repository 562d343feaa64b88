// K8 + K9: causal softmax attention (HF GPT2Attention._attn, stage 2 GPT-2 backbone) and the stage-1
// relative-position attention (optimus_txl_decoder.py:305-387), flash style: the T x T score matrix is
// never written to HBM.  One templated family (REL = false / true):
//
//   forward   grid (q tiles, B*H): online softmax over key tiles, O and (m, l) in registers / smem
//   backward  two kernels without atomics on the big tensors:
//               dkv: grid (k tiles, B*H) loops over the visible q tiles -> dK, dV (+ dr via atomics, REL)
//               dq : grid (q tiles, B*H) loops over the visible k tiles -> dQ (+ d r_w_bias, d r_r_bias)
//
// Scores:  s(i,j) = scale * ( (q_i [+ r_w_bias]) . k_j  [+ (q_i + r_r_bias) . r_{p(i,j)}] ),
//          visible iff j <= i + (Tk - Tq);   REL: p(i,j) = Tq - 1 - i + j  (row of r holding distance
//          i + mlen - j, r row p <-> distance Tk-1-p).  For a (q tile, k tile) pair the needed r rows are
//          one contiguous band of BQ+BK-1 rows, so the BD term is G = (q + r_r_bias) . band^T followed by
//          the skew  BD[ii][jj] = G[ii][BQ-1-ii+jj]  (the reference's _rel_shift).
// Dropout: GPT-2 drops attention PROBABILITIES after the softmax (attn_dropout): the mask scales the
//          P.V product only.  Stage 1 drops and then RENORMALISES (P / (sum P + 1e-8), :362-363), which
//          is a softmax over the randomly kept keys: the mask joins the causal mask inside the softmax.
// Small products run through BlockGemm (mma.sync bf16 tensor cores, or fp32 FMA in the parity mode).
#include "block_gemm.cuh"

constexpr int AE = 64;   // head dim

template <typename T> struct AttnCfg;
template <> struct AttnCfg<bf16> { static constexpr int BQ = 64, BK = 64; };
template <> struct AttnCfg<float> { static constexpr int BQ = 32, BK = 32; };

template <typename T> __device__ __forceinline__ float a_exp(float x);
template <> __device__ __forceinline__ float a_exp<bf16>(float x) { return __expf(x); }
template <> __device__ __forceinline__ float a_exp<float>(float x) { return expf(x); }

struct AttnArgs {
  const void *q, *k, *v, *r, *out, *dout;
  const float *r_w_bias, *r_r_bias, *lse;
  void *o, *dq, *dk, *dv;
  float *lse_out, *dr, *d_rw, *d_rr;
  int64_t ld_q, ld_kv, ld_r, ld_o, ld_dq, ld_dkv;
  int B, Tq, Tk, H;
  float scale, keep_scale;
  uint32_t drop_thr;
  uint64_t seed;
};

// rows [row0, row0+ROWS) x 64 of a [.., ld] matrix -> smem (zero fill outside [0, nrows)), + optional bias[64]
template <typename T, int ROWS, int LD>
__device__ __forceinline__ void a_load(const T* __restrict__ base, int64_t ld, int row0, int nrows, T (*dst)[LD],
                                       const float* __restrict__ bias) {
  constexpr int N = Vec<T>::N, VPR = AE / N;
  for (int i = threadIdx.x; i < ROWS * VPR; i += BG_THREADS) {
    int rr = i / VPR, part = i % VPR, row = row0 + rr;
    Vec<T> t;
    if (row >= 0 && row < nrows) t.load(base + (int64_t)row * ld + part * N);
    else {
#pragma unroll
      for (int j = 0; j < N; ++j) t.v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) dst[rr][part * N + j] = from_f<T>(bias ? t.v[j] + bias[part * N + j] : t.v[j]);
  }
}

template <typename T, int ROWS, int LD>
__device__ __forceinline__ void a_store(T* __restrict__ base, int64_t ld, int row0, int nrows, T (*src)[LD]) {
  constexpr int N = Vec<T>::N, VPR = AE / N;
  for (int i = threadIdx.x; i < ROWS * VPR; i += BG_THREADS) {
    int rr = i / VPR, part = i % VPR, row = row0 + rr;
    if (row < nrows) {
      Vec<T> t;
#pragma unroll
      for (int j = 0; j < N; ++j) t.v[j] = to_f(src[rr][part * N + j]);
      t.store(base + (int64_t)row * ld + part * N);
    }
  }
}

__device__ __forceinline__ bool a_keep(const AttnArgs& a, int bh, int i, int j) {
  if (!a.drop_thr) return true;
  return emo_drop_keep(a.seed, ((uint64_t)bh * a.Tq + i) * (uint64_t)a.Tk + j, a.drop_thr);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, bool REL> struct AttnSmemFwd {
  static constexpr int BQ = AttnCfg<T>::BQ, BK = AttnCfg<T>::BK, BB = REL ? BQ + BK : 16;
  T q[BQ][bg_ld<T>(AE)];
  T qv[REL ? BQ : 1][bg_ld<T>(AE)];
  T k[BK][bg_ld<T>(AE)];
  T v[BK][bg_ld<T>(AE)];
  T rb[REL ? BB : 1][bg_ld<T>(AE)];
  T p[BQ][bg_ld<T>(BK)];
  float s[BQ][BK + 1];
  float g[REL ? BQ : 1][BB + 1];
  float m[BQ], l[BQ], alpha[BQ];
};

template <typename T, bool REL>
__global__ void __launch_bounds__(BG_THREADS) attn_fwd_kernel(const AttnArgs a) {
  using S = AttnSmemFwd<T, REL>;
  constexpr int BQ = S::BQ, BK = S::BK, BB = S::BB;
  constexpr int TPR = BG_THREADS / BQ, CPT = BK / TPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int i0 = blockIdx.x * BQ;
  const int off = a.Tk - a.Tq;
  const T* qg = (const T*)a.q + (int64_t)b * a.Tq * a.ld_q + h * AE;
  const T* kg = (const T*)a.k + (int64_t)b * a.Tk * a.ld_kv + h * AE;
  const T* vg = (const T*)a.v + (int64_t)b * a.Tk * a.ld_kv + h * AE;
  const T* rg = REL ? (const T*)a.r + h * AE : nullptr;

  a_load<T, BQ>(qg, a.ld_q, i0, a.Tq, sm.q, REL ? a.r_w_bias + h * AE : nullptr);
  if constexpr (REL) a_load<T, BQ>(qg, a.ld_q, i0, a.Tq, sm.qv, a.r_r_bias + h * AE);
  for (int i = threadIdx.x; i < BQ; i += BG_THREADS) { sm.m[i] = -INFINITY; sm.l[i] = 0.f; }
  BlockGemm<BQ, AE, T> go;
  go.clear();
  const int jmax = min(a.Tk - 1, i0 + BQ - 1 + off);     // last visible key of this q tile
  const int row = threadIdx.x / TPR, part = threadIdx.x % TPR;

  for (int j0 = 0; j0 <= jmax; j0 += BK) {
    __syncthreads();
    a_load<T, BK>(kg, a.ld_kv, j0, a.Tk, sm.k, nullptr);
    a_load<T, BK>(vg, a.ld_kv, j0, a.Tk, sm.v, nullptr);
    const int pbase = a.Tq - i0 - BQ + j0;
    if constexpr (REL) a_load<T, BB>(rg, a.ld_r, pbase, a.Tk, sm.rb, nullptr);
    __syncthreads();
    {
      BlockGemm<BQ, BK, T> gs;
      gs.clear();
      gs.template mma<true, true>(&sm.q[0][0], bg_ld<T>(AE), &sm.k[0][0], bg_ld<T>(AE), AE);
      gs.foreach ([&](int r_, int c_, float& x) { sm.s[r_][c_] = x; });
    }
    if constexpr (REL) {
      BlockGemm<BQ, BB, T> gg;
      gg.clear();
      gg.template mma<true, true>(&sm.qv[0][0], bg_ld<T>(AE), &sm.rb[0][0], bg_ld<T>(AE), AE);
      gg.foreach ([&](int r_, int c_, float& x) { sm.g[r_][c_] = x; });
    }
    __syncthreads();
    {  // online softmax, TPR adjacent lanes per row
      const int i = i0 + row;
      float x[CPT];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        int jj = part * CPT + c, j = j0 + jj;
        float sc = sm.s[row][jj];
        if constexpr (REL) sc += sm.g[row][BQ - 1 - row + jj];
        sc *= a.scale;
        bool vis = (j <= i + off) && (j < a.Tk) && (i < a.Tq);
        if (REL && vis) vis = a_keep(a, bh, i, j);
        x[c] = vis ? sc : -INFINITY;
        mx = fmaxf(mx, x[c]);
      }
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_old = sm.m[row];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        int jj = part * CPT + c;
        float pv = (x[c] == -INFINITY) ? 0.f : a_exp<T>(x[c] - m_new);
        sum += pv;
        if (!REL && a.drop_thr) pv = a_keep(a, bh, i, j0 + jj) ? pv * a.keep_scale : 0.f;
        sm.p[row][jj] = from_f<T>(pv);
      }
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      __syncwarp();
      if (part == 0) {
        float al = (m_old == -INFINITY) ? 1.f : a_exp<T>(m_old - m_new);   // (m_new == -inf) implies m_old == -inf
        sm.alpha[row] = al;
        sm.l[row] = sm.l[row] * al + sum;
        sm.m[row] = m_new;
      }
    }
    __syncthreads();
    go.foreach ([&](int r_, int c_, float& x) { x *= sm.alpha[r_]; });
    go.template mma<true, false>(&sm.p[0][0], bg_ld<T>(BK), &sm.v[0][0], bg_ld<T>(AE), BK);
  }
  __syncthreads();
  // stage 1: P / (sum P + 1e-8) with sum P = 1 (or 0 when every key was dropped)
  go.foreach ([&](int r_, int c_, float& x) {
    float l = sm.l[r_];
    float den = REL ? l * (1.f + 1e-8f) : l;
    sm.q[r_][c_] = from_f<T>(l > 0.f ? x / den : 0.f);
  });
  __syncthreads();
  a_store<T, BQ>((T*)a.o + (int64_t)b * a.Tq * a.ld_o + h * AE, a.ld_o, i0, a.Tq, sm.q);
  if (a.lse_out)
    for (int i = threadIdx.x; i < BQ; i += BG_THREADS)
      if (i0 + i < a.Tq) a.lse_out[((int64_t)bh) * a.Tq + i0 + i] = sm.l[i] > 0.f ? sm.m[i] + logf(sm.l[i]) : INFINITY;
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <typename T, bool REL> struct AttnSmemBwd {
  static constexpr int BQ = AttnCfg<T>::BQ, BK = AttnCfg<T>::BK, BB = REL ? BQ + BK : 16;
  T q[BQ][bg_ld<T>(AE)];      // q (+ r_w_bias)
  T qv[REL ? BQ : 1][bg_ld<T>(AE)];
  T k[BK][bg_ld<T>(AE)];
  T v[BK][bg_ld<T>(AE)];
  T dout[BQ][bg_ld<T>(AE)];
  T rb[REL ? BB : 1][bg_ld<T>(AE)];
  T p[BQ][bg_ld<T>(BK)];      // (dropped) probabilities, for dV
  T ds[BQ][bg_ld<T>(BK)];
  T dg[REL ? BQ : 1][bg_ld<T>(BB)];
  float g[REL ? BQ : 1][BB + 1];
  float lse[BQ], dsum[BQ];
};

// lse / D = rowsum(dO * O) of the current q tile
template <typename T, typename S>
__device__ __forceinline__ void a_rowstats(const AttnArgs& a, S& sm, int b, int h, int bh, int i0) {
  constexpr int BQ = S::BQ, N = Vec<T>::N, VPR = AE / N;
  const T* og = (const T*)a.out + (int64_t)b * a.Tq * a.ld_o + h * AE;
  const T* dg = (const T*)a.dout + (int64_t)b * a.Tq * a.ld_o + h * AE;
  for (int i = threadIdx.x; i < BQ * VPR; i += BG_THREADS) {   // BQ*VPR is a multiple of 32: whole warps iterate together
    int rr = i / VPR, part = i % VPR, row = i0 + rr;
    float dot = 0.f;
    if (row < a.Tq) {
      Vec<T> o_, d_;
      o_.load(og + (int64_t)row * a.ld_o + part * N);
      d_.load(dg + (int64_t)row * a.ld_o + part * N);
#pragma unroll
      for (int j = 0; j < N; ++j) dot += o_.v[j] * d_.v[j];
    }
#pragma unroll
    for (int o = VPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (part == 0) {
      sm.dsum[rr] = dot;
      sm.lse[rr] = row < a.Tq ? a.lse[(int64_t)bh * a.Tq + row] : INFINITY;
    }
  }
}

// P and dS of one (q tile, k tile) pair into smem (sm.p, sm.ds [, sm.dg]); needs q, qv, k, v, dout, rb, lse, dsum
template <typename T, bool REL, typename S>
__device__ __forceinline__ void a_tile_grads(const AttnArgs& a, S& sm, int bh, int i0, int j0) {
  constexpr int BQ = S::BQ, BK = S::BK, BB = S::BB;
  const int off = a.Tk - a.Tq;
  if constexpr (REL) {
    BlockGemm<BQ, BB, T> gg;
    gg.clear();
    gg.template mma<true, true>(&sm.qv[0][0], bg_ld<T>(AE), &sm.rb[0][0], bg_ld<T>(AE), AE);
    gg.foreach ([&](int r_, int c_, float& x) { sm.g[r_][c_] = x; });
    for (int i = threadIdx.x; i < BQ * BB; i += BG_THREADS) sm.dg[i / BB][i % BB] = from_f<T>(0.f);
    __syncthreads();
  }
  BlockGemm<BQ, BK, T> gs, gp;
  gs.clear();
  gp.clear();
  gs.template mma<true, true>(&sm.q[0][0], bg_ld<T>(AE), &sm.k[0][0], bg_ld<T>(AE), AE);
  gp.template mma<true, true>(&sm.dout[0][0], bg_ld<T>(AE), &sm.v[0][0], bg_ld<T>(AE), AE);
  const float* dp = reinterpret_cast<const float*>(gp.acc);
  int idx = 0;
  gs.foreach ([&](int r_, int c_, float& x) {
    const float dpv = dp[idx++];
    const int i = i0 + r_, j = j0 + c_;
    float sc = x;
    if constexpr (REL) sc += sm.g[r_][BQ - 1 - r_ + c_];
    sc *= a.scale;
    bool vis = (j <= i + off) && (j < a.Tk) && (i < a.Tq);
    bool keep = true;
    if (a.drop_thr && vis) keep = a_keep(a, bh, i, j);
    if (REL && !keep) vis = false;
    float p = vis ? a_exp<T>(sc - sm.lse[r_]) : 0.f;
    float pd = p, dpd = dpv;
    if (!REL && a.drop_thr) { pd = keep ? p * a.keep_scale : 0.f; dpd = keep ? dpv * a.keep_scale : 0.f; }
    float ds = p * (dpd - sm.dsum[r_]) * a.scale;
    sm.p[r_][c_] = from_f<T>(pd);
    sm.ds[r_][c_] = from_f<T>(ds);
    if constexpr (REL) sm.dg[r_][BQ - 1 - r_ + c_] = from_f<T>(ds);
  });
}

template <typename T, bool REL>
__global__ void __launch_bounds__(BG_THREADS) attn_bwd_dkv_kernel(const AttnArgs a) {
  using S = AttnSmemBwd<T, REL>;
  constexpr int BQ = S::BQ, BK = S::BK, BB = S::BB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int j0 = blockIdx.x * BK;
  const int off = a.Tk - a.Tq;
  const T* qg = (const T*)a.q + (int64_t)b * a.Tq * a.ld_q + h * AE;
  const T* kg = (const T*)a.k + (int64_t)b * a.Tk * a.ld_kv + h * AE;
  const T* vg = (const T*)a.v + (int64_t)b * a.Tk * a.ld_kv + h * AE;
  const T* dog = (const T*)a.dout + (int64_t)b * a.Tq * a.ld_o + h * AE;
  const T* rg = REL ? (const T*)a.r + h * AE : nullptr;
  a_load<T, BK>(kg, a.ld_kv, j0, a.Tk, sm.k, nullptr);
  a_load<T, BK>(vg, a.ld_kv, j0, a.Tk, sm.v, nullptr);
  BlockGemm<BK, AE, T> gk, gv;
  gk.clear();
  gv.clear();
  int ifirst = j0 - off;                       // first query that sees key j0
  if (ifirst < 0) ifirst = 0;
  for (int i0 = (ifirst / BQ) * BQ; i0 < a.Tq; i0 += BQ) {
    __syncthreads();
    a_load<T, BQ>(qg, a.ld_q, i0, a.Tq, sm.q, REL ? a.r_w_bias + h * AE : nullptr);
    if constexpr (REL) a_load<T, BQ>(qg, a.ld_q, i0, a.Tq, sm.qv, a.r_r_bias + h * AE);
    a_load<T, BQ>(dog, a.ld_o, i0, a.Tq, sm.dout, nullptr);
    const int pbase = a.Tq - i0 - BQ + j0;
    if constexpr (REL) a_load<T, BB>(rg, a.ld_r, pbase, a.Tk, sm.rb, nullptr);
    a_rowstats<T>(a, sm, b, h, bh, i0);
    __syncthreads();
    a_tile_grads<T, REL>(a, sm, bh, i0, j0);
    __syncthreads();
    gv.template mma<false, false>(&sm.p[0][0], bg_ld<T>(BK), &sm.dout[0][0], bg_ld<T>(AE), BQ);
    gk.template mma<false, false>(&sm.ds[0][0], bg_ld<T>(BK), &sm.q[0][0], bg_ld<T>(AE), BQ);
    if constexpr (REL) {   // dr[pbase + t] += sum_ii dG[ii][t] (q_ii + r_r_bias)
      BlockGemm<BB, AE, T> gr;
      gr.clear();
      gr.template mma<false, false>(&sm.dg[0][0], bg_ld<T>(BB), &sm.qv[0][0], bg_ld<T>(AE), BQ);
      gr.foreach ([&](int t, int c_, float& x) {
        int p = pbase + t;
        if (p >= 0 && p < a.Tk && x != 0.f) atomicAdd(a.dr + ((int64_t)p * a.H + h) * AE + c_, x);
      });
    }
  }
  __syncthreads();
  gk.foreach ([&](int r_, int c_, float& x) { sm.k[r_][c_] = from_f<T>(x); });
  gv.foreach ([&](int r_, int c_, float& x) { sm.v[r_][c_] = from_f<T>(x); });
  __syncthreads();
  a_store<T, BK>((T*)a.dk + (int64_t)b * a.Tk * a.ld_dkv + h * AE, a.ld_dkv, j0, a.Tk, sm.k);
  a_store<T, BK>((T*)a.dv + (int64_t)b * a.Tk * a.ld_dkv + h * AE, a.ld_dkv, j0, a.Tk, sm.v);
}

template <typename T, bool REL>
__global__ void __launch_bounds__(BG_THREADS) attn_bwd_dq_kernel(const AttnArgs a) {
  using S = AttnSmemBwd<T, REL>;
  constexpr int BQ = S::BQ, BK = S::BK, BB = S::BB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int i0 = blockIdx.x * BQ;
  const int off = a.Tk - a.Tq;
  const T* qg = (const T*)a.q + (int64_t)b * a.Tq * a.ld_q + h * AE;
  const T* kg = (const T*)a.k + (int64_t)b * a.Tk * a.ld_kv + h * AE;
  const T* vg = (const T*)a.v + (int64_t)b * a.Tk * a.ld_kv + h * AE;
  const T* dog = (const T*)a.dout + (int64_t)b * a.Tq * a.ld_o + h * AE;
  const T* rg = REL ? (const T*)a.r + h * AE : nullptr;
  a_load<T, BQ>(qg, a.ld_q, i0, a.Tq, sm.q, REL ? a.r_w_bias + h * AE : nullptr);
  if constexpr (REL) a_load<T, BQ>(qg, a.ld_q, i0, a.Tq, sm.qv, a.r_r_bias + h * AE);
  a_load<T, BQ>(dog, a.ld_o, i0, a.Tq, sm.dout, nullptr);
  a_rowstats<T>(a, sm, b, h, bh, i0);
  BlockGemm<BQ, AE, T> gac, gbd;     // content part and position part of dQ (their column sums are the bias grads)
  gac.clear();
  gbd.clear();
  const int jmax = min(a.Tk - 1, i0 + BQ - 1 + off);
  for (int j0 = 0; j0 <= jmax; j0 += BK) {
    __syncthreads();
    a_load<T, BK>(kg, a.ld_kv, j0, a.Tk, sm.k, nullptr);
    a_load<T, BK>(vg, a.ld_kv, j0, a.Tk, sm.v, nullptr);
    const int pbase = a.Tq - i0 - BQ + j0;
    if constexpr (REL) a_load<T, BB>(rg, a.ld_r, pbase, a.Tk, sm.rb, nullptr);
    __syncthreads();
    a_tile_grads<T, REL>(a, sm, bh, i0, j0);
    __syncthreads();
    gac.template mma<true, false>(&sm.ds[0][0], bg_ld<T>(BK), &sm.k[0][0], bg_ld<T>(AE), BK);
    if constexpr (REL) gbd.template mma<true, false>(&sm.dg[0][0], bg_ld<T>(BB), &sm.rb[0][0], bg_ld<T>(AE), BB);
  }
  __syncthreads();
  if constexpr (REL) {
    // bias gradients: column sums over the tile's rows, via smem (reuse sm.g as [2][64] scratch after zeroing)
    float* scratch = &sm.g[0][0];
    for (int i = threadIdx.x; i < 2 * AE; i += BG_THREADS) scratch[i] = 0.f;
    __syncthreads();
    gac.foreach ([&](int r_, int c_, float& x) { if (i0 + r_ < a.Tq) atomicAdd(scratch + c_, x); });
    gbd.foreach ([&](int r_, int c_, float& x) { if (i0 + r_ < a.Tq) atomicAdd(scratch + AE + c_, x); });
    __syncthreads();
    for (int i = threadIdx.x; i < AE; i += BG_THREADS) {
      atomicAdd(a.d_rw + h * AE + i, scratch[i]);
      atomicAdd(a.d_rr + h * AE + i, scratch[AE + i]);
    }
    const float* bd = reinterpret_cast<const float*>(gbd.acc);
    int idx = 0;
    gac.foreach ([&](int r_, int c_, float& x) { sm.q[r_][c_] = from_f<T>(x + bd[idx++]); });
  } else {
    gac.foreach ([&](int r_, int c_, float& x) { sm.q[r_][c_] = from_f<T>(x); });
  }
  __syncthreads();
  a_store<T, BQ>((T*)a.dq + (int64_t)b * a.Tq * a.ld_dq + h * AE, a.ld_dq, i0, a.Tq, sm.q);
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// attn_tc.cu: the tcgen05 + TMA kernels (bf16, no relative positions); EMO_ATTN_TC=0 / emo_attn_set_tc(0) switch back
int emo_attn_tc_enabled();
int emo_attn_fwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out, int64_t ld_o,
                           float* lse, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, cudaStream_t s);
int emo_attn_bwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* out,
                           const void* dout, int64_t ld_o, const float* lse, void* dq, void* dk, void* dv, int64_t ld_dq,
                           int64_t ld_dkv, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, cudaStream_t s);
// relattn_tc.cu: the stage-1 relative-position attention on the same tcgen05 kernels (position scores through HBM)
bool emo_relattn_tc_ok(int B, int Tq, int Tk, int H);
int emo_relattn_fwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* r, int64_t ld_r,
                              const float* r_w_bias, const float* r_r_bias, void* out, int64_t ld_out, float* lse, int B, int Tq, int Tk,
                              int H, float scale, float drop_p, uint64_t seed, cudaStream_t s);
int emo_relattn_bwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* r, int64_t ld_r,
                              const float* r_w_bias, const float* r_r_bias, const void* out, const void* dout, int64_t ld_out,
                              const float* lse, void* dq, void* dk, void* dv, int64_t ld_dq, int64_t ld_dkv, float* dr, float* d_rw,
                              float* d_rr, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, cudaStream_t s);
// a 128-row query tile per CTA: below this many queries (decode steps) the 64-row mma.sync tiles waste less
constexpr int ATTN_TC_MIN_TQ = 64;

template <typename T, bool REL> static int attn_fwd_launch(const AttnArgs& a, cudaStream_t s) {
  constexpr int BQ = AttnCfg<T>::BQ;
  size_t smem = sizeof(AttnSmemFwd<T, REL>);
  EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.Tq + BQ - 1) / BQ, a.B * a.H);
  attn_fwd_kernel<T, REL><<<grid, BG_THREADS, smem, s>>>(a);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
template <typename T, bool REL> static int attn_bwd_launch(const AttnArgs& a, cudaStream_t s) {
  constexpr int BQ = AttnCfg<T>::BQ, BK = AttnCfg<T>::BK;
  size_t smem = sizeof(AttnSmemBwd<T, REL>);
  EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<T, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<T, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 gk((a.Tk + BK - 1) / BK, a.B * a.H), gq((a.Tq + BQ - 1) / BQ, a.B * a.H);
  attn_bwd_dkv_kernel<T, REL><<<gk, BG_THREADS, smem, s>>>(a);
  EMO_LAUNCH_CHECK();
  attn_bwd_dq_kernel<T, REL><<<gq, BG_THREADS, smem, s>>>(a);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

static int attn_check(const char* who, const AttnArgs& a, int dtype, float drop_p) {
  int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE(dtype == EMO_BF16 || dtype == EMO_F32, "%s: bad dtype %d", who, dtype);
  EMO_REQUIRE(a.Tk >= a.Tq && a.Tq >= 0, "%s: need Tk >= Tq (queries are the last Tq positions)", who);
  EMO_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "%s: bad dropout p", who);
  EMO_REQUIRE(al16(a.q) && al16(a.k) && al16(a.v) && (a.r == nullptr || al16(a.r)), "%s: pointers must be 16-byte aligned", who);
  EMO_REQUIRE((a.ld_q * esz) % 16 == 0 && (a.ld_kv * esz) % 16 == 0 && (a.ld_o * esz) % 16 == 0 && (a.ld_r * esz) % 16 == 0,
              "%s: row strides must be 16-byte multiples", who);
  return EMO_OK;
}

static AttnArgs attn_args(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, int B, int Tq, int Tk,
                          int H, float scale, float drop_p, uint64_t seed) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.q = q; a.k = k; a.v = v; a.ld_q = ld_q; a.ld_kv = ld_kv; a.B = B; a.Tq = Tq; a.Tk = Tk; a.H = H;
  a.scale = scale; a.drop_thr = emo_drop_thr(drop_p); a.keep_scale = 1.f / (1.f - drop_p); a.seed = seed;
  return a;
}

extern "C" int emo_attn_fwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out,
                            int64_t ld_out, float* lse, int B, int Tq, int Tk, int H, float scale, float drop_p,
                            uint64_t seed, int dtype, void* stream) {
  AttnArgs a = attn_args(q, k, v, ld_q, ld_kv, B, Tq, Tk, H, scale, drop_p, seed);
  a.o = out; a.ld_o = ld_out; a.lse_out = lse;
  int rc = attn_check("emo_attn_fwd", a, dtype, drop_p);
  if (rc) return rc;
  EMO_REQUIRE(al16(out), "emo_attn_fwd: out must be 16-byte aligned");
  if (B * H == 0 || Tq == 0) return EMO_OK;
  if (dtype == EMO_BF16 && emo_attn_tc_enabled() && Tq >= ATTN_TC_MIN_TQ && scale > 0.f)
    return emo_attn_fwd_tc_launch(q, k, v, ld_q, ld_kv, out, ld_out, lse, B, Tq, Tk, H, scale, drop_p, seed, (cudaStream_t)stream);
  if (dtype == EMO_BF16) return attn_fwd_launch<bf16, false>(a, (cudaStream_t)stream);
  return attn_fwd_launch<float, false>(a, (cudaStream_t)stream);
}

extern "C" int emo_attn_bwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* out,
                            const void* dout, int64_t ld_out, const float* lse, void* dq, void* dk, void* dv,
                            int64_t ld_dq, int64_t ld_dkv, int B, int Tq, int Tk, int H, float scale, float drop_p,
                            uint64_t seed, int dtype, void* stream) {
  AttnArgs a = attn_args(q, k, v, ld_q, ld_kv, B, Tq, Tk, H, scale, drop_p, seed);
  a.out = out; a.dout = dout; a.ld_o = ld_out; a.lse = lse; a.dq = dq; a.dk = dk; a.dv = dv; a.ld_dq = ld_dq; a.ld_dkv = ld_dkv;
  int rc = attn_check("emo_attn_bwd", a, dtype, drop_p);
  if (rc) return rc;
  int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE(al16(out) && al16(dout) && al16(dq) && al16(dk) && al16(dv) && (ld_dq * esz) % 16 == 0 && (ld_dkv * esz) % 16 == 0,
              "emo_attn_bwd: gradient pointers / strides must be 16-byte aligned");
  EMO_REQUIRE(lse != nullptr, "emo_attn_bwd: lse is required");
  if (B * H == 0 || Tq == 0) return EMO_OK;
  if (dtype == EMO_BF16 && emo_attn_tc_enabled() && Tq >= ATTN_TC_MIN_TQ && scale > 0.f)
    return emo_attn_bwd_tc_launch(q, k, v, ld_q, ld_kv, out, dout, ld_out, lse, dq, dk, dv, ld_dq, ld_dkv, B, Tq, Tk, H, scale, drop_p, seed,
                                  (cudaStream_t)stream);
  if (dtype == EMO_BF16) return attn_bwd_launch<bf16, false>(a, (cudaStream_t)stream);
  return attn_bwd_launch<float, false>(a, (cudaStream_t)stream);
}

extern "C" int emo_relattn_fwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* r,
                               int64_t ld_r, const float* r_w_bias, const float* r_r_bias, void* out, int64_t ld_out,
                               float* lse, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed,
                               int dtype, void* stream) {
  AttnArgs a = attn_args(q, k, v, ld_q, ld_kv, B, Tq, Tk, H, scale, drop_p, seed);
  a.r = r; a.ld_r = ld_r; a.r_w_bias = r_w_bias; a.r_r_bias = r_r_bias; a.o = out; a.ld_o = ld_out; a.lse_out = lse;
  int rc = attn_check("emo_relattn_fwd", a, dtype, drop_p);
  if (rc) return rc;
  EMO_REQUIRE(r && r_w_bias && r_r_bias && al16(out), "emo_relattn_fwd: r / biases required, out 16-byte aligned");
  if (B * H == 0 || Tq == 0) return EMO_OK;
  if (dtype == EMO_BF16 && emo_attn_tc_enabled() && scale > 0.f && emo_relattn_tc_ok(B, Tq, Tk, H) && ld_out == (int64_t)H * AE)
    return emo_relattn_fwd_tc_launch(q, k, v, ld_q, ld_kv, r, ld_r, r_w_bias, r_r_bias, out, ld_out, lse, B, Tq, Tk, H, scale, drop_p, seed,
                                     (cudaStream_t)stream);
  if (dtype == EMO_BF16) return attn_fwd_launch<bf16, true>(a, (cudaStream_t)stream);
  return attn_fwd_launch<float, true>(a, (cudaStream_t)stream);
}

extern "C" int emo_relattn_bwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* r,
                               int64_t ld_r, const float* r_w_bias, const float* r_r_bias, const void* out,
                               const void* dout, int64_t ld_out, const float* lse, void* dq, void* dk, void* dv,
                               int64_t ld_dq, int64_t ld_dkv, float* dr, float* d_r_w_bias, float* d_r_r_bias, int B,
                               int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, int dtype, void* stream) {
  AttnArgs a = attn_args(q, k, v, ld_q, ld_kv, B, Tq, Tk, H, scale, drop_p, seed);
  a.r = r; a.ld_r = ld_r; a.r_w_bias = r_w_bias; a.r_r_bias = r_r_bias; a.out = out; a.dout = dout; a.ld_o = ld_out;
  a.lse = lse; a.dq = dq; a.dk = dk; a.dv = dv; a.ld_dq = ld_dq; a.ld_dkv = ld_dkv; a.dr = dr; a.d_rw = d_r_w_bias; a.d_rr = d_r_r_bias;
  int rc = attn_check("emo_relattn_bwd", a, dtype, drop_p);
  if (rc) return rc;
  int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE(al16(out) && al16(dout) && al16(dq) && al16(dk) && al16(dv) && (ld_dq * esz) % 16 == 0 && (ld_dkv * esz) % 16 == 0,
              "emo_relattn_bwd: gradient pointers / strides must be 16-byte aligned");
  EMO_REQUIRE(r && r_w_bias && r_r_bias && lse && dr && d_r_w_bias && d_r_r_bias, "emo_relattn_bwd: null argument");
  if (B * H == 0 || Tq == 0) return EMO_OK;
  if (dtype == EMO_BF16 && emo_attn_tc_enabled() && scale > 0.f && emo_relattn_tc_ok(B, Tq, Tk, H) && ld_out == (int64_t)H * AE)
    return emo_relattn_bwd_tc_launch(q, k, v, ld_q, ld_kv, r, ld_r, r_w_bias, r_r_bias, out, dout, ld_out, lse, dq, dk, dv, ld_dq, ld_dkv, dr,
                                     d_r_w_bias, d_r_r_bias, B, Tq, Tk, H, scale, drop_p, seed, (cudaStream_t)stream);
  if (dtype == EMO_BF16) return attn_bwd_launch<bf16, true>(a, (cudaStream_t)stream);
  return attn_bwd_launch<float, true>(a, (cudaStream_t)stream);
}
