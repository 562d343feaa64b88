// Device code of the FAVOR+ kernels (see favor.cu for the description).  BG_WARPS / FAVOR_NS may be set by the
// includer: a 16-warp build of the backward was measured (+1.6 % on the kernel, nothing on the step: the kernel is
// bound by the shared-memory pipe, not by latency) and dropped; everything lives in a per-includer namespace.
#pragma once
#include "block_gemm.cuh"
#include <stdlib.h>

#ifndef FAVOR_NS
#define FAVOR_NS favor8
#endif
namespace FAVOR_NS {


constexpr int FE = 64;    // head dim
constexpr int FM = 128;   // feature dim (n_dims)
constexpr int FV = 80;    // padded value width: 64 values + ones column + zero pad
constexpr float F_EPS = 1e-6f;
constexpr float F_S2 = 0.125f;                 // softmax_temp = 1/sqrt(64); x is scaled by sqrt(temp)
constexpr float F_HALF_LOG_M = 2.4260151319598084f;   // 0.5 * ln(128)

template <typename T> struct FavorCfg;
template <> struct FavorCfg<bf16> { static constexpr int C = 64; };
template <> struct FavorCfg<float> { static constexpr int C = 16; };   // parity mode: smaller chunks keep the fp32 tiles within 227 KB

// exp of the feature map.  bf16 mode works in base 2: Omega and the row offsets are pre-multiplied by log2(e)
// when they are staged, so phi = ex2.approx(u' - o') is ONE MUFU instruction per feature (the exp was a quarter
// of the forward's instructions); in the backward du . Omega'^T carries the same factor and is multiplied by kInv = ln 2.
// fp32 (parity) mode keeps expf on unscaled operands.
template <typename T> struct FavorMath;
template <> struct FavorMath<bf16> {
  static constexpr float kScale = 1.4426950408889634f;
  static constexpr float kInv = 0.6931471805599453f;     // 1 / kScale
  static __device__ __forceinline__ float ex(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
};
template <> struct FavorMath<float> {
  static constexpr float kScale = 1.f;
  static constexpr float kInv = 1.f;
  static __device__ __forceinline__ float ex(float x) { return expf(x); }
};

// registers -> shared tile: one packed bf16x2 store per adjacent column pair (bf16), scalar stores (fp32)
template <int M, int N, typename T, typename F>
__device__ __forceinline__ void acc_to_smem(BlockGemm<M, N, T>& g, T* dst, int ld, F f) {
  if constexpr (sizeof(T) == 2) {
    g.foreach2([&](int r, int c, float& a, float& b) { st_pair(dst + r * ld + c, f(r, c, a), f(r, c + 1, b)); });
  } else {
    g.foreach ([&](int r, int c, float& a) { dst[r * ld + c] = f(r, c, a); });
  }
}

template <typename T, int C> struct FavorSmemFwd {
  T xq[C][bg_ld<T>(FE)];
  T xk[C][bg_ld<T>(FE)];
  T om[FE][bg_ld<T>(FE)];
  T v[C][bg_ld<T>(FV)];
  T pq[C][bg_ld<T>(FM)];
  T pk[C][bg_ld<T>(FM)];
  T a[C][bg_ld<T>(C)];
  T s[FM][bg_ld<T>(FV)];
  float oq[C], ok[C], den[C];
};

// segment-sum kernels: x = k rows (fwd) or q rows (bwd), w = [v | 1 | 0] (fwd) or G (bwd)
template <typename T, int C> struct FavorSmemSeg {
  T x[2][C][bg_ld<T>(FE)];       // double-buffered (cp.async prefetch of the next chunk)
  T w[2][C][bg_ld<T>(FV)];
  T raw[2][2][C][bg_ld<T>(FE)];  // bwd only: raw out / dout tiles of the next chunk
  T om[FE][bg_ld<T>(FE)];
  T p[C][bg_ld<T>(FM)];
  float den[2][C];
  float off[C];
};

template <typename T, int C> struct FavorSmemBwd {
  T xq[2][C][bg_ld<T>(FE)];      // q, k, v rows: double-buffered (cp.async prefetch of the next chunk)
  T xk[2][C][bg_ld<T>(FE)];
  T v[2][C][bg_ld<T>(FV)];
  T raw[2][C][bg_ld<T>(FE)];     // out, dout rows of the chunk (dead once G is built -> refilled at once)
  T om[FE][bg_ld<T>(FE)];
  T du[C][bg_ld<T>(FE)];         // d(pre-exp) tile; doubles as the staging buffer of the dq/dk/dv stores
  T g[C][bg_ld<T>(FV)];
  T pq[C][bg_ld<T>(FM)];
  T pk[C][bg_ld<T>(FM)];
  T w[C][bg_ld<T>(FM)];
  T a[C][bg_ld<T>(C)];
  T p[C][bg_ld<T>(C)];
  T s[FM][bg_ld<T>(FV)];
  T r[FM][bg_ld<T>(FV)];
  float den[C];
  float oq[C], ok[C], dof[C];
};

static_assert(sizeof(FavorSmemBwd<bf16, FavorCfg<bf16>::C>) <= 227 * 1024, "bwd smem (bf16) exceeds the 227 KB CTA limit");
static_assert(sizeof(FavorSmemBwd<float, FavorCfg<float>::C>) <= 227 * 1024, "bwd smem (fp32) exceeds the 227 KB CTA limit");
static_assert(BG_THREADS % FavorCfg<bf16>::C == 0, "row helpers: a whole number of threads per chunk row");
static_assert(2 * (sizeof(FavorSmemFwd<bf16, FavorCfg<bf16>::C>) + 1024) <= 227 * 1024, "fwd smem: two CTAs per SM");

// ---- asynchronous tile loads (cp.async / LDGSTS): the next chunk's rows stream into shared memory while the
// current chunk is being computed; rows >= valid are zero-filled by the zero-size form of the copy ----
template <typename T> struct CpA { static constexpr int BYTES = sizeof(T) == 2 ? 16 : 4; static constexpr int EL = BYTES / sizeof(T); };
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src, bool pred) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  int n = pred ? BYTES : 0;
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// issue the copies of a [C x 64] tile of rows (token stride ld); `active` = false issues nothing (tail of the loop)
template <typename T, int C, int LD>
__device__ __forceinline__ void issue_rows(const T* __restrict__ src, int64_t ld, int valid, T (*dst)[LD], bool active) {
  if (!active) return;
  constexpr int EL = CpA<T>::EL, VPR = FE / EL;
  for (int i = threadIdx.x; i < C * VPR; i += BG_THREADS) {
    int row = i / VPR, part = i % VPR;
    bool ok = row < valid;
    cp_async<CpA<T>::BYTES>(&dst[row][part * EL], ok ? src + (int64_t)row * ld + part * EL : src, ok);
  }
}

// off[row] = hs * |x_row|^2 + add  from a tile already in smem; TPR adjacent lanes share a row
template <typename T, int C, int LD>
__device__ __forceinline__ void row_offsets(T (*x)[LD], float* off, float hs, float add) {
  constexpr int TPR = BG_THREADS / C, EPT = FE / TPR;
  const int row = threadIdx.x / TPR, part = threadIdx.x % TPR;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < EPT; ++j) { float v = to_f(x[row][part * EPT + j]); ss += v * v; }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (part == 0) off[row] = ss * hs + add;
}

// constant part of V' = [v | 1 | 0..]: the ones column (harmless for rows >= valid: their phi(k) rows are zero)
template <typename T, int C, int LD>
__device__ __forceinline__ void init_ones(T (*v)[LD]) {
  for (int i = threadIdx.x; i < C * (FV - FE); i += BG_THREADS) {
    int row = i / (FV - FE), col = FE + i % (FV - FE);
    v[row][col] = from_f<T>(col == FE ? 1.f : 0.f);
  }
}

// load a [C x 64] tile of rows (token stride ld) into smem, returning per-row sum of squares
// (rows >= valid are zero-filled).  VPR threads share a row.
template <typename T, int C, int LD>
__device__ __forceinline__ void load_rows(const T* __restrict__ src, int64_t ld, int valid, T (*dst)[LD],
                                          float* sumsq /* may be null */, float hs, float add) {
  constexpr int N = Vec<T>::N;
  constexpr int VPR = FE / N;
  for (int i = threadIdx.x; i < C * VPR; i += BG_THREADS) {
    int row = i / VPR, part = i % VPR;
    Vec<T> t;
    if (row < valid) t.load(src + (int64_t)row * ld + part * N);
    else {
#pragma unroll
      for (int j = 0; j < N; ++j) t.v[j] = 0.f;
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) { dst[row][part * N + j] = from_f<T>(t.v[j]); ss += t.v[j] * t.v[j]; }
    if (sumsq) {
#pragma unroll
      for (int o = VPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (part == 0) sumsq[row] = ss * hs + add;
    }
  }
}

template <typename T, int C, int LD>
__device__ __forceinline__ void store_rows(T* __restrict__ dst, int64_t ld, int valid, T (*src)[LD]) {
  constexpr int N = Vec<T>::N;
  constexpr int VPR = FE / N;
  for (int i = threadIdx.x; i < C * VPR; i += BG_THREADS) {
    int row = i / VPR, part = i % VPR;
    if (row < valid) {
      Vec<T> t;
#pragma unroll
      for (int j = 0; j < N; ++j) t.v[j] = to_f(src[row][part * N + j]);
      t.store(dst + (int64_t)row * ld + part * N);
    }
  }
}

template <typename T, int LD>
__device__ __forceinline__ void load_omega(const float* __restrict__ omega, T (*om)[LD]) {
  const float s = 0.35355339059327373f * FavorMath<T>::kScale;   // 64^(-1/4) (and log2 e in bf16 mode) folded into omega
  constexpr int PER = FE * FE / 4 / BG_THREADS;      // float4 per thread (4 with 8 warps, 2 with 16)
  static_assert(FE * FE == 4 * PER * BG_THREADS, "omega tile");
  float4 v[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(omega) + threadIdx.x + u * BG_THREADS);   // all in flight
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    int i = (threadIdx.x + u * BG_THREADS) * 4;
    T* d = &om[i / FE][i % FE];
    d[0] = from_f<T>(v[u].x * s); d[1] = from_f<T>(v[u].y * s); d[2] = from_f<T>(v[u].z * s); d[3] = from_f<T>(v[u].w * s);
  }
}

// U = X . Om_s -> phi rows in smem (rows >= valid zeroed)
template <typename T, int C, int LDX, int LDO, int LDP>
__device__ __forceinline__ void phi_rows(T (*x)[LDX], T (*om)[LDO], const float* off, int valid, T (*phi)[LDP]) {
  BlockGemm<C, FE, T> g;
  g.clear();
  g.template mma_k<true, false, FE>(&x[0][0], LDX, &om[0][0], LDO);
  acc_to_smem(g, &phi[0][0], LDP, [&](int row, int col, float u) { return row < valid ? FavorMath<T>::ex(u - off[row]) : 0.f; });
  acc_to_smem(g, &phi[0][FE], LDP, [&](int row, int col, float u) { return row < valid ? FavorMath<T>::ex(-u - off[row]) : 0.f; });
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BG_THREADS, 2)
favor_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int64_t ld,
                 const float* __restrict__ omega, T* __restrict__ out, int64_t ld_out, float* __restrict__ den_out,
                 const float* __restrict__ state_in, float* __restrict__ state_out,
                 float* __restrict__ seg_states, int nseg, int seg_chunks, int Tlen, int H) {
  constexpr int C = FavorCfg<T>::C;
  using S = FavorSmemFwd<T, C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int bh = blockIdx.x / nseg, seg = blockIdx.x % nseg;
  const int b = bh / H, h = bh % H;
  const int64_t base = (int64_t)b * Tlen * ld + (int64_t)h * FE;
  const int64_t obase = (int64_t)b * Tlen * ld_out + (int64_t)h * FE;

  load_omega<T>(omega, sm.om);
  BlockGemm<FM, FV, T> gs;   // running prefix state S' (fp32 master)
  gs.clear();
  if (state_in) {            // continue a sequence (decode: append a block of tokens to a running state)
    const float* si = state_in + (int64_t)bh * FM * FV;
    gs.foreach ([&](int row, int col, float& x) { x = si[row * FV + col]; });
  }
  if (seg_states && seg > 0) {   // exclusive prefix over the earlier segments (favor_segsum_kernel + favor_prefix_kernel)
    const float* si = seg_states + ((int64_t)bh * (nseg + 1) + seg) * FM * FV;
    gs.foreach ([&](int row, int col, float& x) { x += si[row * FV + col]; });
  }
  for (int i = threadIdx.x; i < FM * bg_ld<T>(FV); i += BG_THREADS) (&sm.s[0][0])[i] = from_f<T>(0.f);
  __syncthreads();
  gs.foreach ([&](int row, int col, float& x) { sm.s[row][col] = from_f<T>(x); });
  __syncthreads();

  const int t_begin = seg * seg_chunks * C;
  const int t_end = (t_begin + seg_chunks * C < Tlen) ? t_begin + seg_chunks * C : Tlen;
  init_ones<T, C>(sm.v);
  T (*stage)[bg_ld<T>(FE)] = reinterpret_cast<T (*)[bg_ld<T>(FE)]>(&sm.pq[0][0]);   // output staging (pq is dead by then)
  // software pipeline over chunks: group 2c = (q,k) of chunk c, group 2c+1 = v of chunk c
  {
    const int valid0 = (Tlen - t_begin < C) ? (Tlen - t_begin) : C;
    issue_rows<T, C>(q + base + (int64_t)t_begin * ld, ld, valid0, sm.xq, t_begin < t_end);
    issue_rows<T, C>(k + base + (int64_t)t_begin * ld, ld, valid0, sm.xk, t_begin < t_end);
    cp_commit();
    issue_rows<T, C>(v + base + (int64_t)t_begin * ld, ld, valid0, sm.v, t_begin < t_end);
    cp_commit();
  }
  for (int t0 = t_begin; t0 < t_end; t0 += C) {
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    const int tn = t0 + C;
    const bool more = tn < t_end;
    const int validn = (Tlen - tn < C) ? (Tlen - tn) : C;
    cp_wait<1>();            // (q,k) of this chunk have landed (v may still be in flight)
    __syncthreads();
    row_offsets<T, C>(sm.xq, sm.oq, 0.5f * F_S2 * FavorMath<T>::kScale, F_HALF_LOG_M * FavorMath<T>::kScale);
    row_offsets<T, C>(sm.xk, sm.ok, 0.5f * F_S2 * FavorMath<T>::kScale, F_HALF_LOG_M * FavorMath<T>::kScale);
    __syncthreads();
    phi_rows<T, C>(sm.xq, sm.om, sm.oq, valid, sm.pq);
    phi_rows<T, C>(sm.xk, sm.om, sm.ok, valid, sm.pk);
    __syncthreads();
    issue_rows<T, C>(q + base + (int64_t)tn * ld, ld, validn, sm.xq, more);     // prefetch the next chunk's q, k
    issue_rows<T, C>(k + base + (int64_t)tn * ld, ld, validn, sm.xk, more);
    cp_commit();
    {
      BlockGemm<C, C, T> ga;
      ga.clear();
      ga.template mma_k<true, true, FM>(&sm.pq[0][0], bg_ld<T>(FM), &sm.pk[0][0], bg_ld<T>(FM));
      acc_to_smem(ga, &sm.a[0][0], bg_ld<T>(C), [](int row, int col, float x) { return col <= row ? x : 0.f; });
    }
    cp_wait<1>();            // v of this chunk has landed
    __syncthreads();
    {
      BlockGemm<C, FV, T> go;
      go.clear();
      go.template mma_k<true, false, C>(&sm.a[0][0], bg_ld<T>(C), &sm.v[0][0], bg_ld<T>(FV));
      go.template mma_k<true, false, FM>(&sm.pq[0][0], bg_ld<T>(FM), &sm.s[0][0], bg_ld<T>(FV));
      go.foreach ([&](int row, int col, float& x) { if (col == FE) { sm.den[row] = x + F_EPS; sm.oq[row] = 1.f / (x + F_EPS); } });
      __syncthreads();     // oq (the phi(q) offsets) is dead here: it carries 1/den to the output scaling
      if constexpr (sizeof(T) == 2) {
        go.foreach2([&](int row, int col, float& x0, float& x1) { if (col < FE) { float inv = sm.oq[row]; st_pair(&stage[row][col], x0 * inv, x1 * inv); } });
      } else {
        go.foreach ([&](int row, int col, float& x) { if (col < FE) stage[row][col] = x * sm.oq[row]; });
      }
    }
    __syncthreads();
    store_rows<T, C>(out + obase + (int64_t)t0 * ld_out, ld_out, valid, stage);
    if (den_out)
      for (int i = threadIdx.x; i < valid; i += BG_THREADS) den_out[((int64_t)b * Tlen + t0 + i) * H + h] = sm.den[i];
    gs.template mma_k<false, false, C>(&sm.pk[0][0], bg_ld<T>(FM), &sm.v[0][0], bg_ld<T>(FV));
    __syncthreads();         // every warp is done reading v and s
    acc_to_smem(gs, &sm.s[0][0], bg_ld<T>(FV), [](int, int, float x) { return x; });
    issue_rows<T, C>(v + base + (int64_t)tn * ld, ld, validn, sm.v, more);      // prefetch the next chunk's v
    cp_commit();
  }
  cp_wait<0>();
  if (seg == nseg - 1) {       // the last segment ends on the final prefix state
    if (state_out) {
      float* so = state_out + (int64_t)bh * FM * FV;
      gs.foreach ([&](int row, int col, float& x) { so[row * FV + col] = x; });
    }
    if (seg_states) {          // ... which is also the last slot of the workspace (what the backward starts from)
      float* so = seg_states + ((int64_t)bh * (nseg + 1) + nseg) * FM * FV;
      gs.foreach ([&](int row, int col, float& x) { so[row * FV + col] = x; });
    }
  }
  if (seg_states && seg == 0) {  // slot 0 = the empty prefix (kept defined for readers of the workspace)
    float* so = seg_states + (int64_t)bh * (nseg + 1) * FM * FV;
    for (int i = threadIdx.x; i < FM * FV; i += BG_THREADS) so[i] = 0.f;
  }
}

// segment-local sums of the prefix state: seg_states[bh][seg] = sum_{t in segment} phi(k_t)^T [v_t | 1 | 0]
template <typename T>
__global__ void __launch_bounds__(BG_THREADS, 2)
favor_segsum_kernel(const T* __restrict__ k, const T* __restrict__ v, int64_t ld, const float* __restrict__ omega,
                    float* __restrict__ seg_states, int nseg, int seg_chunks, int Tlen, int H) {
  constexpr int C = FavorCfg<T>::C;
  using S = FavorSmemSeg<T, C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  // one CTA per (b, h, segment) for every segment but the last (whose sum no later segment needs): the local
  // sum of segment s lands in slot s + 1, where the in-place scan turns it into the prefix of segment s + 1
  const int bh = blockIdx.x / (nseg - 1), seg = blockIdx.x % (nseg - 1);
  const int b = bh / H, h = bh % H;
  const int64_t base = (int64_t)b * Tlen * ld + (int64_t)h * FE;
  load_omega<T>(omega, sm.om);
  init_ones<T, C>(sm.w[0]);
  init_ones<T, C>(sm.w[1]);
  BlockGemm<FM, FV, T> gs;
  gs.clear();
  const int t_begin = seg * seg_chunks * C;
  const int t_end = (t_begin + seg_chunks * C < Tlen) ? t_begin + seg_chunks * C : Tlen;
  {
    const int valid0 = (Tlen - t_begin < C) ? (Tlen - t_begin) : C;
    issue_rows<T, C>(k + base + (int64_t)t_begin * ld, ld, valid0, sm.x[0], t_begin < t_end);
    issue_rows<T, C>(v + base + (int64_t)t_begin * ld, ld, valid0, sm.w[0], t_begin < t_end);
    cp_commit();
  }
  int buf = 0;
  for (int t0 = t_begin; t0 < t_end; t0 += C, buf ^= 1) {
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    const int tn = t0 + C;
    const int validn = (Tlen - tn < C) ? (Tlen - tn) : C;
    __syncthreads();         // the other buffer is free (its mma of the previous iteration is done)
    issue_rows<T, C>(k + base + (int64_t)tn * ld, ld, validn, sm.x[buf ^ 1], tn < t_end);
    issue_rows<T, C>(v + base + (int64_t)tn * ld, ld, validn, sm.w[buf ^ 1], tn < t_end);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    row_offsets<T, C>(sm.x[buf], sm.off, 0.5f * F_S2 * FavorMath<T>::kScale, F_HALF_LOG_M * FavorMath<T>::kScale);
    __syncthreads();
    phi_rows<T, C>(sm.x[buf], sm.om, sm.off, valid, sm.p);
    __syncthreads();
    gs.template mma_k<false, false, C>(&sm.p[0][0], bg_ld<T>(FM), &sm.w[buf][0][0], bg_ld<T>(FV));
  }
  cp_wait<0>();
  float* so = seg_states + ((int64_t)bh * (nseg + 1) + seg + 1) * FM * FV;
  gs.foreach ([&](int row, int col, float& x) { so[row * FV + col] = x; });
}

// In-place scan over the slots of every (b,h): forward -> slot s = sum of the local sums of segments < s (slot 0 and
// slot nseg, the total, are written by the main kernel); reverse -> slot s = sum of the local sums of segments > s.
__global__ void favor_prefix_kernel(float* __restrict__ states, int nseg, int reverse, int64_t n_bh) {
  constexpr int64_t per4 = (int64_t)FM * FV / 4;          // float4 per slot
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bh * per4) return;
  float4* p = reinterpret_cast<float4*>(states) + (i / per4) * (nseg + 1) * per4 + (i % per4);
  float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
  auto add = [](float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; };
  if (!reverse) {            // slots 1..nseg-1 hold the local sums of segments 0..nseg-2 -> running (inclusive) sums
    for (int s = 1; s < nseg; ++s) {
      add(run, p[s * per4]);
      p[s * per4] = run;
    }
  } else {
    for (int s = nseg - 1; s >= 0; --s) {
      float4 t = p[s * per4];
      p[s * per4] = run;
      add(run, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward (single reverse pass)
// ---------------------------------------------------------------------------------------------
// dphi (already multiplied by phi, in sm.w) -> du (smem, T) and do (per row, fp32)
template <typename T, int C, typename SM>
__device__ __forceinline__ void phi_bwd_reduce(SM& sm) {
  constexpr int TPR = BG_THREADS / C;     // threads per row
  constexpr int FPT = FE / TPR;           // features per thread
  int row = threadIdx.x / TPR, part = threadIdx.x % TPR;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < FPT; ++j) {
    int f = part * FPT + j;
    float wp = to_f(sm.w[row][f]), wm = to_f(sm.w[row][FE + f]);
    sm.du[row][f] = from_f<T>(wp - wm);
    acc += wp + wm;
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (part == 0) sm.dof[row] = -acc;
}

// G = [dout/den | -(dout.out)/den | 0] for one chunk, from the raw out / dout tiles in smem (rows >= valid -> 0)
template <typename T, int C, int LD, int LDG>
__device__ __forceinline__ void make_g(T (*o_)[LD], T (*d_)[LD], const float* den, int valid, T (*g)[LDG]) {
  constexpr int TPR = BG_THREADS / C, EPT = FE / TPR;
  const int row = threadIdx.x / TPR, part = threadIdx.x % TPR;
  const float inv = row < valid ? 1.f / den[row] : 0.f;
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    float ov = to_f(o_[row][part * EPT + j]), dv = to_f(d_[row][part * EPT + j]);
    dot += ov * dv;
    g[row][part * EPT + j] = from_f<T>(dv * inv);
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (part == 0) {
    g[row][FE] = from_f<T>(-dot * inv);
#pragma unroll
    for (int j = FE + 1; j < FV; ++j) g[row][j] = from_f<T>(0.f);
  }
}

// den of one chunk: C floats, one 4-byte async copy per row
template <int C>
__device__ __forceinline__ void issue_den(const float* __restrict__ den, int stride, int valid, float* dst, bool active) {
  if (!active) return;
  for (int i = threadIdx.x; i < C; i += BG_THREADS) cp_async<4>(&dst[i], i < valid ? den + (int64_t)i * stride : den, i < valid);
}

// segment-local sums of the reverse state: seg_rstates[bh][seg] = sum_{t in segment} phi(q_t)^T G_t
template <typename T>
__global__ void __launch_bounds__(BG_THREADS, 2)
favor_bwd_segsum_kernel(const T* __restrict__ q, int64_t ld, const float* __restrict__ omega, const T* __restrict__ out,
                        const T* __restrict__ dout, int64_t ld_out, const float* __restrict__ den_in,
                        float* __restrict__ seg_rstates, int nseg, int seg_chunks, int Tlen, int H) {
  constexpr int C = FavorCfg<T>::C;
  using S = FavorSmemSeg<T, C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int bh = blockIdx.x / nseg, seg = blockIdx.x % nseg;
  const int b = bh / H, h = bh % H;
  const int64_t base = (int64_t)b * Tlen * ld + (int64_t)h * FE;
  const int64_t obase = (int64_t)b * Tlen * ld_out + (int64_t)h * FE;
  load_omega<T>(omega, sm.om);
  BlockGemm<FM, FV, T> gr;
  gr.clear();
  const int t_begin = seg * seg_chunks * C;
  const int t_end = (t_begin + seg_chunks * C < Tlen) ? t_begin + seg_chunks * C : Tlen;
  auto issue = [&](int t0, int buf) {
    const bool act = t0 < t_end;
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    issue_rows<T, C>(q + base + (int64_t)t0 * ld, ld, valid, sm.x[buf], act);
    issue_rows<T, C>(out + obase + (int64_t)t0 * ld_out, ld_out, valid, sm.raw[buf][0], act);
    issue_rows<T, C>(dout + obase + (int64_t)t0 * ld_out, ld_out, valid, sm.raw[buf][1], act);
    issue_den<C>(den_in + ((int64_t)b * Tlen + t0) * H + h, H, valid, sm.den[buf], act);
    cp_commit();
  };
  issue(t_begin, 0);
  int buf = 0;
  for (int t0 = t_begin; t0 < t_end; t0 += C, buf ^= 1) {
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    __syncthreads();         // the other buffers are free
    issue(t0 + C, buf ^ 1);
    cp_wait<1>();
    __syncthreads();
    row_offsets<T, C>(sm.x[buf], sm.off, 0.5f * F_S2 * FavorMath<T>::kScale, F_HALF_LOG_M * FavorMath<T>::kScale);
    make_g<T, C>(sm.raw[buf][0], sm.raw[buf][1], sm.den[buf], valid, sm.w[0]);
    __syncthreads();
    phi_rows<T, C>(sm.x[buf], sm.om, sm.off, valid, sm.p);
    __syncthreads();
    gr.template mma_k<false, false, C>(&sm.p[0][0], bg_ld<T>(FM), &sm.w[0][0][0], bg_ld<T>(FV));
  }
  cp_wait<0>();
  float* so = seg_rstates + ((int64_t)bh * (nseg + 1) + seg) * FM * FV;
  gr.foreach ([&](int row, int col, float& x) { so[row * FV + col] = x; });
}

template <typename T>
__global__ void __launch_bounds__(BG_THREADS, 1)
favor_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int64_t ld,
                 const float* __restrict__ omega, const T* __restrict__ out, const T* __restrict__ dout,
                 int64_t ld_out, const float* __restrict__ den_in, const float* __restrict__ seg_states,
                 const float* __restrict__ seg_rstates, int nseg, int seg_chunks, int fwd_nseg, int ratio,
                 T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv, int64_t ld_d, int Tlen, int H) {
  constexpr int C = FavorCfg<T>::C;
  using S = FavorSmemBwd<T, C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int bh = blockIdx.x / nseg, seg = blockIdx.x % nseg;
  const int b = bh / H, h = bh % H;
  const int64_t base = (int64_t)b * Tlen * ld + (int64_t)h * FE;
  const int64_t obase = (int64_t)b * Tlen * ld_out + (int64_t)h * FE;
  const int64_t dbase = (int64_t)b * Tlen * ld_d + (int64_t)h * FE;
  const int nchunk = (Tlen + C - 1) / C;
  const int c_begin = seg * seg_chunks;
  const int c_end = (c_begin + seg_chunks < nchunk) ? c_begin + seg_chunks : nchunk;

  // chunk loads: one cp.async group per chunk; q/k/v into buffer `buf`, out/dout/den into the single raw set
  auto issue_qkv = [&](int c, int buf) {
    if (c < c_begin) return;
    const int t0 = c * C;
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    issue_rows<T, C>(q + base + (int64_t)t0 * ld, ld, valid, sm.xq[buf], true);
    issue_rows<T, C>(k + base + (int64_t)t0 * ld, ld, valid, sm.xk[buf], true);
    issue_rows<T, C>(v + base + (int64_t)t0 * ld, ld, valid, sm.v[buf], true);
  };
  auto issue_raw = [&](int c) {
    if (c < c_begin) return;
    const int t0 = c * C;
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    issue_rows<T, C>(out + obase + (int64_t)t0 * ld_out, ld_out, valid, sm.raw[0], true);
    issue_rows<T, C>(dout + obase + (int64_t)t0 * ld_out, ld_out, valid, sm.raw[1], true);
    issue_den<C>(den_in + ((int64_t)b * Tlen + t0) * H + h, H, valid, sm.den, true);
  };
  issue_qkv(c_end - 1, 0);
  issue_raw(c_end - 1);
  cp_commit();

  load_omega<T>(omega, sm.om);
  init_ones<T, C>(sm.v[0]);
  init_ones<T, C>(sm.v[1]);
  for (int i = threadIdx.x; i < FM * bg_ld<T>(FV); i += BG_THREADS) (&sm.r[0][0])[i] = from_f<T>(0.f);
  BlockGemm<FM, FV, T> gs, gr;   // forward prefix state (rolled back) and reverse state
  gs.clear();
  gr.clear();
  {   // prefix state at the END of this segment = the forward's exclusive prefix at that boundary (last slot = total)
    int slot = (seg + 1) * ratio;
    if (slot > fwd_nseg) slot = fwd_nseg;
    const float* si = seg_states + ((int64_t)bh * (fwd_nseg + 1) + slot) * FM * FV;
    gs.foreach ([&](int row, int col, float& x) { x = si[row * FV + col]; });
  }
  if (nseg > 1) {   // reverse state carried in from the later segments (suffix-exclusive, favor_prefix_kernel)
    const float* si = seg_rstates + ((int64_t)bh * (nseg + 1) + seg) * FM * FV;
    gr.foreach ([&](int row, int col, float& x) { x = si[row * FV + col]; });
  }
  __syncthreads();
  gr.foreach ([&](int row, int col, float& x) { sm.r[row][col] = from_f<T>(x); });

  int buf = 0;
  for (int c = c_end - 1; c >= c_begin; --c, buf ^= 1) {
    const int t0 = c * C;
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    T (*xq)[bg_ld<T>(FE)] = sm.xq[buf];
    T (*xk)[bg_ld<T>(FE)] = sm.xk[buf];
    T (*vv)[bg_ld<T>(FV)] = sm.v[buf];
    cp_wait<0>();            // this chunk's q, k, v, out, dout, den have landed
    __syncthreads();
    row_offsets<T, C>(xq, sm.oq, 0.5f * F_S2 * FavorMath<T>::kScale, F_HALF_LOG_M * FavorMath<T>::kScale);
    row_offsets<T, C>(xk, sm.ok, 0.5f * F_S2 * FavorMath<T>::kScale, F_HALF_LOG_M * FavorMath<T>::kScale);
    make_g<T, C>(sm.raw[0], sm.raw[1], sm.den, valid, sm.g);
    __syncthreads();
    issue_qkv(c - 1, buf ^ 1);     // prefetch the previous chunk (reverse order) while this one is computed
    issue_raw(c - 1);
    cp_commit();
    phi_rows<T, C>(xq, sm.om, sm.oq, valid, sm.pq);
    phi_rows<T, C>(xk, sm.om, sm.ok, valid, sm.pk);
    __syncthreads();
    {  // roll the prefix state back to the start of this chunk
      BlockGemm<FM, FV, T> tmp;
      tmp.clear();
      tmp.template mma_k<false, false, C>(&sm.pk[0][0], bg_ld<T>(FM), &vv[0][0], bg_ld<T>(FV));
      constexpr int NA = sizeof(gs.acc) / sizeof(float);
      float* a = reinterpret_cast<float*>(gs.acc);
      const float* t = reinterpret_cast<const float*>(tmp.acc);
#pragma unroll
      for (int i = 0; i < NA; ++i) a[i] -= t[i];
      acc_to_smem(gs, &sm.s[0][0], bg_ld<T>(FV), [](int, int, float x) { return x; });
    }
    {
      BlockGemm<C, C, T> ga;
      ga.clear();
      ga.template mma_k<true, true, FM>(&sm.pq[0][0], bg_ld<T>(FM), &sm.pk[0][0], bg_ld<T>(FM));
      acc_to_smem(ga, &sm.a[0][0], bg_ld<T>(C), [](int row, int col, float x) { return col <= row ? x : 0.f; });
      ga.clear();
      ga.template mma_k<true, true, FV>(&sm.g[0][0], bg_ld<T>(FV), &vv[0][0], bg_ld<T>(FV));
      acc_to_smem(ga, &sm.p[0][0], bg_ld<T>(C), [](int row, int col, float x) { return col <= row ? x : 0.f; });
    }
    __syncthreads();
    // ---- dq ----
    {
      BlockGemm<C, FM, T> gd;
      gd.clear();
      gd.template mma_k<true, false, C>(&sm.p[0][0], bg_ld<T>(C), &sm.pk[0][0], bg_ld<T>(FM));
      gd.template mma_k<true, true, FV>(&sm.g[0][0], bg_ld<T>(FV), &sm.s[0][0], bg_ld<T>(FV));
      acc_to_smem(gd, &sm.w[0][0], bg_ld<T>(FM), [&](int row, int col, float x) { return x * to_f(sm.pq[row][col]); });
    }
    __syncthreads();
    phi_bwd_reduce<T, C>(sm);
    __syncthreads();
    {
      BlockGemm<C, FE, T> gx;
      gx.clear();
      gx.template mma_k<true, true, FE>(&sm.du[0][0], bg_ld<T>(FE), &sm.om[0][0], bg_ld<T>(FE));
      __syncthreads();       // du is re-used as the store staging buffer
      acc_to_smem(gx, &sm.du[0][0], bg_ld<T>(FE), [&](int row, int col, float x) {
        return x * FavorMath<T>::kInv + sm.dof[row] * F_S2 * to_f(xq[row][col]);
      });
    }
    __syncthreads();
    store_rows<T, C>(dq + dbase + (int64_t)t0 * ld_d, ld_d, valid, sm.du);
    // ---- dk ----
    {
      BlockGemm<C, FM, T> gd;
      gd.clear();
      gd.template mma_k<false, false, C>(&sm.p[0][0], bg_ld<T>(C), &sm.pq[0][0], bg_ld<T>(FM));
      gd.template mma_k<true, true, FV>(&vv[0][0], bg_ld<T>(FV), &sm.r[0][0], bg_ld<T>(FV));
      acc_to_smem(gd, &sm.w[0][0], bg_ld<T>(FM), [&](int row, int col, float x) { return x * to_f(sm.pk[row][col]); });
    }
    __syncthreads();         // also: the dq store has finished reading du
    phi_bwd_reduce<T, C>(sm);
    __syncthreads();
    {
      BlockGemm<C, FE, T> gx;
      gx.clear();
      gx.template mma_k<true, true, FE>(&sm.du[0][0], bg_ld<T>(FE), &sm.om[0][0], bg_ld<T>(FE));
      __syncthreads();
      acc_to_smem(gx, &sm.du[0][0], bg_ld<T>(FE), [&](int row, int col, float x) {
        return x * FavorMath<T>::kInv + sm.dof[row] * F_S2 * to_f(xk[row][col]);
      });
    }
    __syncthreads();
    store_rows<T, C>(dk + dbase + (int64_t)t0 * ld_d, ld_d, valid, sm.du);
    __syncthreads();
    // ---- dv ----
    {
      BlockGemm<C, FE, T> gv;      // only the 64 value columns of A^T G + Phi_k R are needed
      gv.clear();
      gv.template mma_k<false, false, C>(&sm.a[0][0], bg_ld<T>(C), &sm.g[0][0], bg_ld<T>(FV));
      gv.template mma_k<true, false, FM>(&sm.pk[0][0], bg_ld<T>(FM), &sm.r[0][0], bg_ld<T>(FV));
      acc_to_smem(gv, &sm.du[0][0], bg_ld<T>(FE), [](int, int, float x) { return x; });
    }
    __syncthreads();
    store_rows<T, C>(dv + dbase + (int64_t)t0 * ld_d, ld_d, valid, sm.du);
    // ---- reverse state ----
    gr.template mma_k<false, false, C>(&sm.pq[0][0], bg_ld<T>(FM), &sm.g[0][0], bg_ld<T>(FV));
    __syncthreads();         // every warp is done with r (dk, dv) and with du (dv store)
    acc_to_smem(gr, &sm.r[0][0], bg_ld<T>(FV), [](int, int, float x) { return x; });
  }
  cp_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// decode step (recurrent form): one CTA of 256 threads per (sequence, head).  The [128 x 80] fp32 state is
// streamed once (read + write, 16-byte vectors, consecutive threads on consecutive columns of a row);
// out = phi(q) . S' normalised by its ones-column.
// ---------------------------------------------------------------------------------------------
constexpr int FS_CG = 17;                    // float4 column groups that carry data: 64 values + the ones column
constexpr int FS_RL = 15;                    // row lanes: 17 x 15 = 255 threads stream the state
template <typename T>
__global__ void __launch_bounds__(256) favor_step_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                         const T* __restrict__ v, int64_t ld,
                                                         const float* __restrict__ omega, float* __restrict__ state,
                                                         T* __restrict__ out, int64_t ld_out, int H) {
  __shared__ float xq[FE], xk[FE], pq[FM], pk[FM];
  __shared__ __align__(16) float vv[FS_CG * 4];
  __shared__ float red[FS_RL][FS_CG * 4];
  __shared__ __align__(16) float om_s[FE * FE];
  const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
  const float s = 0.35355339059327373f;
  {  // Omega (constant) -> shared memory with one batch of coalesced 16-byte loads (a 64-step chain of strided global
     // loads per feature was most of this kernel's 8 us)
    float4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = __ldg(reinterpret_cast<const float4*>(omega) + tid + 256 * u);
#pragma unroll
    for (int u = 0; u < 4; ++u) reinterpret_cast<float4*>(om_s)[tid + 256 * u] = t[u];
  }
  pdl_trigger();
  pdl_wait();
  // the state rows of this thread: all 9 loads in flight now, consumed after the feature map is ready
  constexpr int NR = (FM + FS_RL - 1) / FS_RL;
  float4 st[NR];
  if (tid < FS_CG * FS_RL) {
    const float* sb0 = state + (int64_t)blockIdx.x * FM * FV + (tid % FS_CG) * 4;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int f = tid / FS_CG + i * FS_RL;
      if (f < FM) st[i] = *reinterpret_cast<const float4*>(sb0 + f * FV);
    }
  }
  if (tid < FE) {
    xq[tid] = to_f(q[(int64_t)b * ld + h * FE + tid]) * s;
    xk[tid] = to_f(k[(int64_t)b * ld + h * FE + tid]) * s;
    vv[tid] = to_f(v[(int64_t)b * ld + h * FE + tid]);
  } else if (tid < FS_CG * 4) {
    vv[tid] = (tid == FE) ? 1.f : 0.f;
  }
  __syncthreads();
  if (tid < 2 * FE) {  // thread f<64 -> q feature f ; thread 64+f -> k feature f
    const float* x = (tid < FE) ? xq : xk;
    int f = tid & (FE - 1);
    float u = 0.f, n2 = 0.f;
#pragma unroll 8
    for (int e = 0; e < FE; ++e) { u = fmaf(x[e], om_s[e * FE + f], u); n2 = fmaf(x[e], x[e], n2); }
    float o = 0.5f * n2 + F_HALF_LOG_M;
    float* p = (tid < FE) ? pq : pk;
    p[f] = expf(u - o);
    p[FE + f] = expf(-u - o);
  }
  __syncthreads();
  if (tid < FS_CG * FS_RL) {
    const int cg = tid % FS_CG, rl = tid / FS_CG;
    const float4 v4 = *reinterpret_cast<const float4*>(&vv[cg * 4]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float* sbase = state + (int64_t)blockIdx.x * FM * FV + cg * 4;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int f = rl + i * FS_RL;
      if (f >= FM) break;
      float4* sp = reinterpret_cast<float4*>(sbase + f * FV);
      float4 sv = st[i];
      const float a = pk[f], c = pq[f];
      sv.x = fmaf(a, v4.x, sv.x); sv.y = fmaf(a, v4.y, sv.y); sv.z = fmaf(a, v4.z, sv.z); sv.w = fmaf(a, v4.w, sv.w);
      *sp = sv;
      acc.x = fmaf(c, sv.x, acc.x); acc.y = fmaf(c, sv.y, acc.y); acc.z = fmaf(c, sv.z, acc.z); acc.w = fmaf(c, sv.w, acc.w);
    }
    *reinterpret_cast<float4*>(&red[rl][cg * 4]) = acc;
  }
  __syncthreads();
  if (tid <= FE) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < FS_RL; ++r) t += red[r][tid];
    red[0][tid] = t;
  }
  __syncthreads();
  if (tid < FE) out[(int64_t)b * ld_out + h * FE + tid] = from_f<T>(red[0][tid] / (red[0][FE] + F_EPS));
}


}  // namespace FAVOR_NS
