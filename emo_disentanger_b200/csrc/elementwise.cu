// Memory-bound kernels of the hot path: embedding gather, LayerNorm fwd/bwd, dropout, column sums,
// cross-entropy, grad-norm + Adam, casts.  All HBM-bound: 16-byte vector accesses, coalesced rows,
// grids sized in multiples of the SM count where the kernel is persistent.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void emo_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* emo_last_error(void) { return g_err; }
extern "C" int emo_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// K1 embedding forward: out = dropout((E_tok[tok] + E_seg[seg]) * scale + pe[t])
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void embed_fwd_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ seg,
                                 int64_t sb, int64_t st, const float* __restrict__ e_tok,
                                 const float* __restrict__ e_seg, const float* __restrict__ pe,
                                 T* __restrict__ out, int B, int T_, int d, float scale,
                                 uint32_t thr, float keep_scale, uint64_t seed) {
  constexpr int N = Vec<T>::N;
  const int vec_per_row = d / N;
  int64_t total = (int64_t)B * T_ * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / vec_per_row;
    int col = (int)(i % vec_per_row) * N;
    int b = (int)(row / T_), t = (int)(row % T_);
    int64_t id = tok[b * sb + t * st];
    const float* e = e_tok + id * d + col;
    const float* s = (seg != nullptr && e_seg != nullptr) ? e_seg + seg[b * sb + t * st] * d + col : nullptr;
    const float* p = pe ? pe + (int64_t)t * d + col : nullptr;
    Vec<T> o;
#pragma unroll
    for (int j = 0; j < N; j += 4) {
      float4 ev = *reinterpret_cast<const float4*>(e + j);
      float4 sv = s ? *reinterpret_cast<const float4*>(s + j) : make_float4(0, 0, 0, 0);
      float4 pv = p ? *reinterpret_cast<const float4*>(p + j) : make_float4(0, 0, 0, 0);
      // reference order: tok*scale, then += seg*scale, then + pe
      o.v[j + 0] = (ev.x * scale + sv.x * scale) + pv.x;
      o.v[j + 1] = (ev.y * scale + sv.y * scale) + pv.y;
      o.v[j + 2] = (ev.z * scale + sv.z * scale) + pv.z;
      o.v[j + 3] = (ev.w * scale + sv.w * scale) + pv.w;
    }
    if (thr) {
      uint64_t e0 = (uint64_t)row * d + col;
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        uint32_t h = emo_drop_hash(seed, e0 + j);
        o.v[j] = ((h & 0xffffu) >= thr) ? o.v[j] * keep_scale : 0.f;
        o.v[j + 1] = ((h >> 16) >= thr) ? o.v[j + 1] * keep_scale : 0.f;
      }
    }
    o.store(out + row * d + col);
  }
}

extern "C" int emo_embed_fwd(const int64_t* tok, const int64_t* seg, int64_t stride_b, int64_t stride_t,
                             const float* e_tok, const float* e_seg, const float* pe, void* out, int B,
                             int T, int d, float scale, float drop_p, uint64_t seed, int out_dtype,
                             void* stream) {
  EMO_REQUIRE(d % 8 == 0, "emo_embed_fwd: d must be a multiple of 8");
  EMO_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "emo_embed_fwd: bad dropout p");
  if ((int64_t)B * T == 0) return EMO_OK;
  uint32_t thr = emo_drop_thr(drop_p);
  float ks = 1.f / (1.f - drop_p);
  int64_t total = (int64_t)B * T * d / (out_dtype == EMO_BF16 ? 8 : 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > emo_num_sms() * 16) blocks = emo_num_sms() * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == EMO_BF16)
    embed_fwd_kernel<bf16><<<blocks, 256, 0, s>>>(tok, seg, stride_b, stride_t, e_tok, e_seg, pe, (bf16*)out, B, T, d, scale, thr, ks, seed);
  else
    embed_fwd_kernel<float><<<blocks, 256, 0, s>>>(tok, seg, stride_b, stride_t, e_tok, e_seg, pe, (float*)out, B, T, d, scale, thr, ks, seed);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// decode-time embedding: one row per sequence, token / segment / POSITION read from device memory, so a
// ragged batch of independent generations is one static-shaped (CUDA-graph capturable) launch.
template <typename T>
__global__ void embed_rows_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ seg,
                                  const int64_t* __restrict__ pos, const float* __restrict__ e_tok,
                                  const float* __restrict__ e_seg, const float* __restrict__ pe,
                                  T* __restrict__ out, int rows, int d, float scale, int64_t* __restrict__ pos_advance) {
  pdl_trigger();
  pdl_wait();
  int row = blockIdx.x;
  if (row >= rows) return;
  const float* e = e_tok + tok[row] * d;
  const float* s = (seg && e_seg) ? e_seg + seg[row] * d : nullptr;
  const int64_t pr = pos ? pos[row] : 0;
  const float* p = (pos && pe) ? pe + pr * d : nullptr;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = e[c] * scale;
    if (s) v += s[c] * scale;
    if (p) v += p[c];
    out[(int64_t)row * d + c] = from_f<T>(v);
  }
  if (pos_advance) {          // the sequence moves on by one position (every thread has read pos[row] by now)
    __syncthreads();
    if (threadIdx.x == 0) pos_advance[row] = pr + 1;
  }
}
static int g_emo_pdl = 0;
int emo_pdl_enabled() { return g_emo_pdl; }
extern "C" void emo_set_pdl(int on) { g_emo_pdl = on; }
extern "C" int emo_embed_rows(const int64_t* tok, const int64_t* seg, const int64_t* pos, const float* e_tok,
                              const float* e_seg, const float* pe, void* out, int rows, int d, float scale,
                              int64_t* pos_advance, int out_dtype, void* stream) {
  if (rows == 0) return EMO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == EMO_BF16)
    EMO_CHECK_CUDA(emo_launch_dep(embed_rows_kernel<bf16>, dim3(rows), dim3(128), 0, s, tok, seg, pos, e_tok, e_seg, pe, (bf16*)out, rows, d, scale, pos_advance));
  else
    EMO_CHECK_CUDA(emo_launch_dep(embed_rows_kernel<float>, dim3(rows), dim3(128), 0, s, tok, seg, pos, e_tok, e_seg, pe, (float*)out, rows, d, scale, pos_advance));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// embedding backward: each block owns ROWS consecutive tokens; 128 threads x 4 columns (d = 512).
// Segment-table gradient is reduced in registers per block (2 rows), token table via atomics.
template <typename T>
__global__ void embed_bwd_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ seg,
                                 int64_t sb, int64_t st, const T* __restrict__ dout,
                                 float* __restrict__ d_e_tok, float* __restrict__ d_e_seg, int B, int T_,
                                 int d, float scale, uint32_t thr, float keep_scale, uint64_t seed,
                                 int64_t pad_idx, int rows_per_block) {
  int64_t rows = (int64_t)B * T_;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (int col = threadIdx.x * 4; col < d; col += blockDim.x * 4) {
    float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int64_t row = r0; row < r1; ++row) {
      int b = (int)(row / T_), t = (int)(row % T_);
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = to_f(dout[row * d + col + j]) * scale;
      if (thr) {
        uint64_t e0 = (uint64_t)row * d + col;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          uint32_t h = emo_drop_hash(seed, e0 + j);
          g[j] = ((h & 0xffffu) >= thr) ? g[j] * keep_scale : 0.f;
          g[j + 1] = ((h >> 16) >= thr) ? g[j + 1] * keep_scale : 0.f;
        }
      }
      int64_t id = tok[b * sb + t * st];
      if (id != pad_idx)      // one 16-byte reduction per thread and row (red.global.add.v4.f32) instead of four scalar ones
        atomicAdd(reinterpret_cast<float4*>(d_e_tok + id * d + col), make_float4(g[0], g[1], g[2], g[3]));
      if (seg != nullptr && d_e_seg != nullptr) {
        int sidx = (int)seg[b * sb + t * st] & 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[sidx][j] += g[j];
      }
    }
    if (seg != nullptr && d_e_seg != nullptr) {
#pragma unroll
      for (int sidx = 0; sidx < 2; ++sidx)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(d_e_seg + sidx * d + col + j, acc[sidx][j]);
    }
  }
}

extern "C" int emo_embed_bwd(const int64_t* tok, const int64_t* seg, int64_t stride_b, int64_t stride_t,
                             const void* dout, float* d_e_tok, float* d_e_seg, int B, int T, int d,
                             float scale, float drop_p, uint64_t seed, int64_t pad_idx, int dtype,
                             void* stream) {
  EMO_REQUIRE(d % 4 == 0, "emo_embed_bwd: d must be a multiple of 4");
  int64_t rows = (int64_t)B * T;
  if (rows == 0) return EMO_OK;
  uint32_t thr = emo_drop_thr(drop_p);
  float ks = 1.f / (1.f - drop_p);
  int rpb = 32;
  int blocks = (int)((rows + rpb - 1) / rpb);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == EMO_BF16)
    embed_bwd_kernel<bf16><<<blocks, 128, 0, s>>>(tok, seg, stride_b, stride_t, (const bf16*)dout, d_e_tok, d_e_seg, B, T, d, scale, thr, ks, seed, pad_idx, rpb);
  else
    embed_bwd_kernel<float><<<blocks, 128, 0, s>>>(tok, seg, stride_b, stride_t, (const float*)dout, d_e_tok, d_e_seg, B, T, d, scale, thr, ks, seed, pad_idx, rpb);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// ---------------------------------------------------------------------------------------------
// K2 LayerNorm, d = 512, one warp per row, row held in registers (16 values per lane)
// ---------------------------------------------------------------------------------------------
constexpr int LN_D = 512;

template <typename T> struct RowRegs {
  static constexpr int N = Vec<T>::N;
  static constexpr int NV = LN_D / (32 * N);
  float v[NV * N];
  __device__ __forceinline__ static int col(int j, int lane) { return (j * 32 + lane) * N; }
  __device__ __forceinline__ void load(const T* row, int lane) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      Vec<T> t;
      t.load(row + col(j, lane));
#pragma unroll
      for (int i = 0; i < N; ++i) v[j * N + i] = t.v[i];
    }
  }
  __device__ __forceinline__ void store(T* row, int lane) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      Vec<T> t;
#pragma unroll
      for (int i = 0; i < N; ++i) t.v[i] = v[j * N + i];
      t.store(row + col(j, lane));
    }
  }
};

// RES: the row is x + res, summed in fp32 (x = the projection's output WITHOUT its residual, res = the sub-layer's
// input).  The pre-LN sum then never exists as a rounded 16-bit tensor on the forward path -- at 12 layers that
// rounding is a quarter of the bf16 hidden-state error -- and sum_out (optional) keeps a copy for the backward.
template <typename T, bool RES>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                     T* __restrict__ sum_out, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                     int64_t rows, float eps) {
  // persistent warps: gamma / beta stay in registers; each warp streams rows with the loads of the next DEPTH rows
  // already in flight, kept as packed 16-byte vectors (memory-bound: 2 x 1 KiB per row at bf16 -- bytes in flight
  // per SM, not arithmetic, set the achieved bandwidth)
  using R = RowRegs<T>;
  constexpr int E = R::NV * R::N;
  constexpr int DEPTH = 3;
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float g_[E], b_[E];
#pragma unroll
  for (int j = 0; j < R::NV; ++j)
#pragma unroll
    for (int i = 0; i < R::N; i += 4) {
      float4 g4 = *reinterpret_cast<const float4*>(gamma + R::col(j, lane) + i);
      float4 b4 = *reinterpret_cast<const float4*>(beta + R::col(j, lane) + i);
      g_[j * R::N + i] = g4.x; g_[j * R::N + i + 1] = g4.y; g_[j * R::N + i + 2] = g4.z; g_[j * R::N + i + 3] = g4.w;
      b_[j * R::N + i] = b4.x; b_[j * R::N + i + 1] = b4.y; b_[j * R::N + i + 2] = b4.z; b_[j * R::N + i + 3] = b4.w;
    }
  constexpr int NS = RES ? 2 : 1;
  uint4 q[DEPTH][NS][R::NV];              // ring of raw rows: q[k] = row + k * wstride
  auto fetch = [&](uint4 (&dst)[NS][R::NV], int64_t r) {
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < R::NV; ++j) dst[0][j] = *reinterpret_cast<const uint4*>(x + r * LN_D + R::col(j, lane));
      if constexpr (RES) {
#pragma unroll
        for (int j = 0; j < R::NV; ++j) dst[1][j] = *reinterpret_cast<const uint4*>(res + r * LN_D + R::col(j, lane));
      }
    }
  };
  auto unpack = [&](const uint4 (&src)[R::NV], R& r) {
#pragma unroll
    for (int j = 0; j < R::NV; ++j) {
      if constexpr (sizeof(T) == 2) {
        unpack_bf16x2(src[j].x, r.v[j * 8], r.v[j * 8 + 1]); unpack_bf16x2(src[j].y, r.v[j * 8 + 2], r.v[j * 8 + 3]);
        unpack_bf16x2(src[j].z, r.v[j * 8 + 4], r.v[j * 8 + 5]); unpack_bf16x2(src[j].w, r.v[j * 8 + 6], r.v[j * 8 + 7]);
      } else {
        r.v[j * 4] = __uint_as_float(src[j].x); r.v[j * 4 + 1] = __uint_as_float(src[j].y);
        r.v[j * 4 + 2] = __uint_as_float(src[j].z); r.v[j * 4 + 3] = __uint_as_float(src[j].w);
      }
    }
  };
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) fetch(q[k], row + k * wstride);
  for (; row < rows; row += wstride) {
    R r;
    unpack(q[0][0], r);
    if constexpr (RES) {
      R r2;
      unpack(q[0][1], r2);
#pragma unroll
      for (int i = 0; i < E; ++i) r.v[i] += r2.v[i];
      if (sum_out) r.store(sum_out + row * LN_D, lane);
    }
#pragma unroll
    for (int k = 0; k + 1 < DEPTH; ++k)
#pragma unroll
      for (int s_ = 0; s_ < NS; ++s_)
#pragma unroll
        for (int j = 0; j < R::NV; ++j) q[k][s_][j] = q[k + 1][s_][j];
    fetch(q[DEPTH - 1], row + DEPTH * wstride);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) s += r.v[i];
    float mu = warp_sum(s) * (1.f / LN_D);
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) { float dlt = r.v[i] - mu; qq += dlt * dlt; }
    float rs = rsqrtf(warp_sum(qq) * (1.f / LN_D) + eps);
#pragma unroll
    for (int i = 0; i < E; ++i) r.v[i] = (r.v[i] - mu) * rs * g_[i] + b_[i];
    r.store(y + row * LN_D, lane);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
  }
}

static int ln_fwd_launch(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta, void* y,
                         float* mean, float* rstd, int64_t rows, float eps, int dtype, cudaStream_t s) {
  int64_t want = (rows + 7) / 8;
  int blocks = (int)(want < (int64_t)emo_num_sms() * 8 ? want : (int64_t)emo_num_sms() * 8);
  if (dtype == EMO_BF16) {
    if (res) ln_fwd_kernel<bf16, true><<<blocks, 256, 0, s>>>((const bf16*)x, (const bf16*)res, (bf16*)sum_out, gamma, beta, (bf16*)y, mean, rstd, rows, eps);
    else ln_fwd_kernel<bf16, false><<<blocks, 256, 0, s>>>((const bf16*)x, nullptr, nullptr, gamma, beta, (bf16*)y, mean, rstd, rows, eps);
  } else {
    if (res) ln_fwd_kernel<float, true><<<blocks, 256, 0, s>>>((const float*)x, (const float*)res, (float*)sum_out, gamma, beta, (float*)y, mean, rstd, rows, eps);
    else ln_fwd_kernel<float, false><<<blocks, 256, 0, s>>>((const float*)x, nullptr, nullptr, gamma, beta, (float*)y, mean, rstd, rows, eps);
  }
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

extern "C" int emo_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                          float* rstd, int64_t rows, int d, float eps, int dtype, void* stream) {
  EMO_REQUIRE(d == LN_D, "emo_ln_fwd: d must be 512 (got %d)", d);
  if (rows == 0) return EMO_OK;
  return ln_fwd_launch(x, nullptr, nullptr, gamma, beta, y, mean, rstd, rows, eps, dtype, (cudaStream_t)stream);
}

extern "C" int emo_ln_res_fwd(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta,
                              void* y, float* mean, float* rstd, int64_t rows, int d, float eps, int dtype,
                              void* stream) {
  EMO_REQUIRE(d == LN_D, "emo_ln_res_fwd: d must be 512 (got %d)", d);
  EMO_REQUIRE(res != nullptr, "emo_ln_res_fwd: res is required (use emo_ln_fwd without a residual)");
  if (rows == 0) return EMO_OK;
  return ln_fwd_launch(x, res, sum_out, gamma, beta, y, mean, rstd, rows, eps, dtype, (cudaStream_t)stream);
}

template <typename T, int DEPTH, int NTH, int MINB>
__global__ void __launch_bounds__(NTH, MINB) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const T* __restrict__ add_in,
                                                     T* __restrict__ dx, T* __restrict__ dx_drop, uint32_t thr,
                                                     float keep_scale, uint64_t seed, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, float* __restrict__ dxsum,
                                                     int64_t rows) {
  using R = RowRegs<T>;
  constexpr int E = R::NV * R::N;
  __shared__ float s_dg[LN_D], s_db[LN_D], s_dx[LN_D];
  for (int i = threadIdx.x; i < LN_D; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; s_dx[i] = 0.f; }
  __syncthreads();
  int lane = threadIdx.x & 31;
  int wpb = blockDim.x >> 5;
  float g_[E], dg[E], db[E], dxs[E];
#pragma unroll
  for (int j = 0; j < R::NV; ++j)
#pragma unroll
    for (int i = 0; i < R::N; ++i) { g_[j * R::N + i] = gamma[R::col(j, lane) + i]; dg[j * R::N + i] = 0.f; db[j * R::N + i] = 0.f; dxs[j * R::N + i] = 0.f; }
  // The next row of every input is requested (as packed 16-byte vectors) before this row is reduced: with 8-16 warps
  // per SM at 128 registers, one row per warp in flight left the kernel at half of the HBM roofline (latency x
  // bandwidth needs ~35 KB in flight per SM).  Rows are private to a warp, so dx may still alias dy / add_in.
  uint4 qd[DEPTH][R::NV], qx[DEPTH][R::NV], qa[DEPTH][R::NV];
  float mu_n[DEPTH], rs_n[DEPTH];
  auto fetch = [&](int k, int64_t r) {
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < R::NV; ++j) {
        qd[k][j] = *reinterpret_cast<const uint4*>(dy + r * LN_D + R::col(j, lane));
        qx[k][j] = *reinterpret_cast<const uint4*>(x + r * LN_D + R::col(j, lane));
        if (add_in) qa[k][j] = *reinterpret_cast<const uint4*>(add_in + r * LN_D + R::col(j, lane));
      }
      mu_n[k] = mean[r];
      rs_n[k] = rstd[r];
    }
  };
  auto unpack = [&](const uint4 (&src)[R::NV], R& r) {
#pragma unroll
    for (int j = 0; j < R::NV; ++j) {
      if constexpr (sizeof(T) == 2) {
        unpack_bf16x2(src[j].x, r.v[j * 8], r.v[j * 8 + 1]); unpack_bf16x2(src[j].y, r.v[j * 8 + 2], r.v[j * 8 + 3]);
        unpack_bf16x2(src[j].z, r.v[j * 8 + 4], r.v[j * 8 + 5]); unpack_bf16x2(src[j].w, r.v[j * 8 + 6], r.v[j * 8 + 7]);
      } else {
        r.v[j * 4] = __uint_as_float(src[j].x); r.v[j * 4 + 1] = __uint_as_float(src[j].y);
        r.v[j * 4 + 2] = __uint_as_float(src[j].z); r.v[j * 4 + 3] = __uint_as_float(src[j].w);
      }
    }
  };
  const int64_t wstride = (int64_t)gridDim.x * wpb;
  int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) fetch(k, row + k * wstride);
  for (; row < rows; row += wstride) {
    R rdy, rx;
    uint4 ca[R::NV];                       // add_in of this row stays packed until it is added
    unpack(qd[0], rdy);
    unpack(qx[0], rx);
#pragma unroll
    for (int j = 0; j < R::NV; ++j) ca[j] = qa[0][j];
    const float mu = mu_n[0], rs = rs_n[0];
#pragma unroll
    for (int k = 0; k + 1 < DEPTH; ++k) {
#pragma unroll
      for (int j = 0; j < R::NV; ++j) { qd[k][j] = qd[k + 1][j]; qx[k][j] = qx[k + 1][j]; qa[k][j] = qa[k + 1][j]; }
      mu_n[k] = mu_n[k + 1];
      rs_n[k] = rs_n[k + 1];
    }
    fetch(DEPTH - 1, row + DEPTH * wstride);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      float xh = (rx.v[i] - mu) * rs;
      float g = g_[i] * rdy.v[i];
      dg[i] += rdy.v[i] * xh;
      db[i] += rdy.v[i];
      c1 += g;
      c2 += g * xh;
      rx.v[i] = xh;
      rdy.v[i] = g;
    }
    c1 = warp_sum(c1) * (1.f / LN_D);
    c2 = warp_sum(c2) * (1.f / LN_D);
#pragma unroll
    for (int i = 0; i < E; ++i) rdy.v[i] = rs * (rdy.v[i] - c1 - rx.v[i] * c2);
    if (add_in) {
      R ra;
      unpack(ca, ra);
#pragma unroll
      for (int i = 0; i < E; ++i) rdy.v[i] += ra.v[i];
    }
    rdy.store(dx + row * LN_D, lane);
    if (dx_drop) {
      if (thr) {
#pragma unroll
        for (int j = 0; j < R::NV; ++j) {
          uint64_t e0 = (uint64_t)row * LN_D + R::col(j, lane);
#pragma unroll
          for (int i = 0; i < R::N; i += 2) {
            uint32_t h = emo_drop_hash(seed, e0 + i);
            float& a = rdy.v[j * R::N + i];
            float& b = rdy.v[j * R::N + i + 1];
            a = ((h & 0xffffu) >= thr) ? a * keep_scale : 0.f;
            b = ((h >> 16) >= thr) ? b * keep_scale : 0.f;
          }
        }
      }
      rdy.store(dx_drop + row * LN_D, lane);
    }
    if (dxsum) {      // column sums of the tensor that feeds the projection below (dx_drop if written, else dx)
#pragma unroll
      for (int i = 0; i < E; ++i) dxs[i] += to_f(from_f<T>(rdy.v[i]));
    }
  }
#pragma unroll
  for (int j = 0; j < R::NV; ++j)
#pragma unroll
    for (int i = 0; i < R::N; ++i) {
      atomicAdd(&s_dg[R::col(j, lane) + i], dg[j * R::N + i]);
      atomicAdd(&s_db[R::col(j, lane) + i], db[j * R::N + i]);
      if (dxsum) atomicAdd(&s_dx[R::col(j, lane) + i], dxs[j * R::N + i]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < LN_D; i += blockDim.x) {
    atomicAdd(dgamma + i, s_dg[i]);
    atomicAdd(dbeta + i, s_db[i]);
    if (dxsum) atomicAdd(dxsum + i, s_dx[i]);
  }
}

extern "C" int emo_ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                          const float* gamma, const void* add_in, void* dx, void* dx_drop, float drop_p,
                          uint64_t seed, float* dgamma, float* dbeta, float* dxsum, int64_t rows, int d, int dtype,
                          void* stream) {
  EMO_REQUIRE(d == LN_D, "emo_ln_bwd: d must be 512 (got %d)", d);
  if (rows == 0) return EMO_OK;
  uint32_t thr = emo_drop_thr(drop_p);
  float ks = 1.f / (1.f - drop_p);
  cudaStream_t s = (cudaStream_t)stream;
  // persistent, one wave: 3 resident CTAs of 4 warps per SM at 168 registers (no spills), the next row of every input
  // in flight per warp.  Measured at 151 552 rows (scripts/ln_perf.py): no prefetch, 2 x 8 warps 165-175 us;
  // 1 x 8 warps with 3 rows in flight 146 us; 3 x 4 warps with 2 rows in flight 157 us (spills); this one 130 us.
  const int64_t want = (rows + 3) / 4;
  int blocks = (int)(want < (int64_t)emo_num_sms() * 3 ? want : (int64_t)emo_num_sms() * 3);
#define EMO_LN_BWD(T, D, N, B) ln_bwd_kernel<T, D, N, B><<<blocks, N, 0, s>>>((const T*)dy, (const T*)x, mean, rstd, gamma, (const T*)add_in, (T*)dx, (T*)dx_drop, thr, ks, seed, dgamma, dbeta, dxsum, rows)
  if (dtype == EMO_BF16) EMO_LN_BWD(bf16, 1, 128, 3);
  else EMO_LN_BWD(float, 1, 128, 1);
#undef EMO_LN_BWD
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// ---------------------------------------------------------------------------------------------
// elementwise dropout (same hash as everywhere else; element index = flat index)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, uint32_t thr,
                               float keep_scale, uint64_t seed) {
  constexpr int N = Vec<T>::N;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * N; i < n; i += (int64_t)gridDim.x * blockDim.x * N) {
    Vec<T> t;
    t.load(x + i);
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      uint32_t h = emo_drop_hash(seed, (uint64_t)i + j);
      t.v[j] = ((h & 0xffffu) >= thr) ? t.v[j] * keep_scale : 0.f;
      t.v[j + 1] = ((h >> 16) >= thr) ? t.v[j + 1] * keep_scale : 0.f;
    }
    t.store(y + i);
  }
}
extern "C" int emo_dropout_apply(const void* x, void* y, int64_t n, float drop_p, uint64_t seed, int dtype,
                                 void* stream) {
  EMO_REQUIRE(n % 8 == 0, "emo_dropout_apply: n must be a multiple of 8");
  if (n == 0) return EMO_OK;
  uint32_t thr = emo_drop_thr(drop_p);
  float ks = 1.f / (1.f - drop_p);
  int64_t vecs = n / (dtype == EMO_BF16 ? 8 : 4);
  int blocks = (int)((vecs + 255) / 256);
  if (blocks > emo_num_sms() * 16) blocks = emo_num_sms() * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == EMO_BF16) dropout_kernel<bf16><<<blocks, 256, 0, s>>>((const bf16*)x, (bf16*)y, n, thr, ks, seed);
  else dropout_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, (float*)y, n, thr, ks, seed);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients).  Vector kernel: a warp covers 32 x 16 B of one row (256 bf16 / 128 fp32
// columns, whole 128-byte lines), 8 warps x 4 rows in flight per CTA; fp32 partials meet in smem, one
// atomic per column per CTA.  Scalar kernel: unaligned / odd shapes.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, int64_t ld, int64_t M, int64_t N,
                                                         int64_t Npad, float* __restrict__ out, int rows_per_block) {
  constexpr int VN = Vec<T>::N;
  constexpr int CW = 32 * VN;                 // columns per CTA
  __shared__ float s[8][CW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * CW + lane * VN;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  float acc[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) acc[j] = 0.f;
  if (c < Npad) {                             // Npad % VN == 0: a vector is all-in or all-out (pad columns are read, not summed out)
    int64_t r = r0 + w;
    for (; r + 24 < r1; r += 32) {            // 4 independent 16-byte loads in flight per thread
      Vec<T> t0, t1, t2, t3;
      t0.load(x + r * ld + c);
      t1.load(x + (r + 8) * ld + c);
      t2.load(x + (r + 16) * ld + c);
      t3.load(x + (r + 24) * ld + c);
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[j] += (t0.v[j] + t1.v[j]) + (t2.v[j] + t3.v[j]);
    }
    for (; r < r1; r += 8) {
      Vec<T> t;
      t.load(x + r * ld + c);
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[j] += t.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) s[w][lane * VN + j] = acc[j];
  __syncthreads();
  for (int i = threadIdx.x; i < CW; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][i];
    int64_t cc = (int64_t)blockIdx.x * CW + i;
    if (cc < N) atomicAdd(out + cc, t);
  }
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t M, int64_t N, float* __restrict__ out,
                              int rows_per_block) {
  __shared__ float s[8][64];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t c = (int64_t)blockIdx.x * 64 + lane * 2;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float a0 = 0.f, a1 = 0.f;
  if (c < N) {
    bool two = (c + 1 < N);
    for (int64_t r = r0 + w; r < r1; r += 8) {
      a0 += to_f(x[r * ld + c]);
      if (two) a1 += to_f(x[r * ld + c + 1]);
    }
  }
  s[w][lane * 2] = a0;
  s[w][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
    int64_t cc = (int64_t)blockIdx.x * 64 + threadIdx.x;
    if (cc < N) atomicAdd(out + cc, t);
  }
}
extern "C" int emo_colsum(const void* x, int64_t ld, int64_t M, int64_t N, float* out, int dtype, void* stream) {
  if (M == 0 || N == 0) return EMO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int vn = dtype == EMO_BF16 ? 8 : 4;
  // ragged N is fine when the row is padded (ld >= N rounded up to a vector): the pad columns are read, not summed
  const int64_t npad = (N + vn - 1) / vn * vn;
  const bool vec = (ld % vn == 0) && (npad <= ld) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    const int cw = 32 * vn;
    const unsigned gx = (unsigned)((npad + cw - 1) / cw);
    int64_t want = (4 * (int64_t)emo_num_sms() + gx - 1) / gx;          // ~4 CTAs per SM
    int64_t rpb = (M + want - 1) / want;
    rpb = (rpb + 31) / 32 * 32;
    if (rpb < 32) rpb = 32;
    dim3 grid(gx, (unsigned)((M + rpb - 1) / rpb));
    if (dtype == EMO_BF16) colsum_vec_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, ld, M, N, npad, out, (int)rpb);
    else colsum_vec_kernel<float><<<grid, 256, 0, s>>>((const float*)x, ld, M, N, npad, out, (int)rpb);
    EMO_LAUNCH_CHECK();
    return EMO_OK;
  }
  int rpb = 256;
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + rpb - 1) / rpb));
  if (dtype == EMO_BF16) colsum_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, ld, M, N, out, rpb);
  else colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)x, ld, M, N, out, rpb);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// ---------------------------------------------------------------------------------------------
// K10 cross-entropy: one warp per row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t tgt_at(const int64_t* tgt, int64_t r, int64_t inner, int64_t so, int64_t si) {
  return tgt[(r / inner) * so + (r % inner) * si];
}

__global__ void ce_count_kernel(const int64_t* __restrict__ tgt, int64_t rows, int64_t inner, int64_t so,
                                int64_t si, int64_t ignore, float* __restrict__ count) {
  float c = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
    c += (tgt_at(tgt, r, inner, so, si) != ignore) ? 1.f : 0.f;
  c = warp_sum(c);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    if (t != 0.f) atomicAdd(count, t);
  }
}
extern "C" int emo_ce_count(const int64_t* tgt, int64_t rows, int64_t tgt_inner, int64_t tgt_stride_outer,
                            int64_t tgt_stride_inner, int64_t ignore_index, float* count, void* stream) {
  if (rows == 0) return EMO_OK;
  int blocks = (int)((rows + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  ce_count_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tgt, rows, tgt_inner, tgt_stride_outer, tgt_stride_inner, ignore_index, count);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

template <typename TD>
__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ logits, int64_t ld,
                                                 const int64_t* __restrict__ tgt, int64_t rows, int64_t inner,
                                                 int64_t so, int64_t si, int V, int64_t ignore,
                                                 const float* __restrict__ count, float gscale,
                                                 float* __restrict__ loss_sum, float* __restrict__ ncorrect,
                                                 int32_t* __restrict__ pred, TD* __restrict__ dl, int64_t ld_dl) {
  __shared__ float s_loss[8], s_corr[8];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float my_loss = 0.f, my_corr = 0.f;
  float inv_count = 0.f;
  if (dl) { float c = count[0]; inv_count = c > 0.f ? gscale / c : 0.f; }
  for (int64_t r = (int64_t)blockIdx.x * wpb + w; r < rows; r += (int64_t)gridDim.x * wpb) {
    const float* row = logits + r * ld;
    int64_t t = tgt_at(tgt, r, inner, so, si);
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
      float v = row[c];
      if (v > mx) { mx = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (ov > mx || (ov == mx && oa < am)) { mx = ov; am = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < V; c += 32) se += __expf(row[c] - mx);
    se = warp_sum(se);
    bool valid = (t != ignore);
    if (pred && lane == 0) pred[r] = am;
    if (valid && lane == 0) {
      my_loss += (logf(se) + mx) - row[t];
      my_corr += (am == (int)t) ? 1.f : 0.f;
    }
    if (dl) {
      TD* drow = dl + r * ld_dl;
      float inv_se = 1.f / se;
      for (int c = lane; c < (int)ld_dl; c += 32) {
        float g = 0.f;
        if (valid && c < V) g = (__expf(row[c] - mx) * inv_se - (c == (int)t ? 1.f : 0.f)) * inv_count;
        drow[c] = from_f<TD>(g);
      }
    }
  }
  if (lane == 0) { s_loss[w] = my_loss; s_corr[w] = my_corr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < wpb; ++i) { a += s_loss[i]; b += s_corr[i]; }
    if (loss_sum) atomicAdd(loss_sum, a);
    if (ncorrect) atomicAdd(ncorrect, b);
  }
}
extern "C" int emo_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* tgt, int64_t rows,
                              int64_t tgt_inner, int64_t tgt_stride_outer, int64_t tgt_stride_inner, int V,
                              int64_t ignore_index, const float* count, float gscale, float* loss_sum,
                              float* ncorrect, int32_t* pred, void* dlogits, int64_t ld_dl, int dl_dtype,
                              void* stream) {
  EMO_REQUIRE(V > 0 && ld >= V, "emo_ce_fwd_bwd: bad V/ld");
  EMO_REQUIRE(dlogits == nullptr || (count != nullptr && ld_dl >= V), "emo_ce_fwd_bwd: dlogits needs count and ld_dl >= V");
  if (rows == 0) return EMO_OK;
  int64_t want = (rows + 7) / 8;
  int blocks = (int)(want < (int64_t)emo_num_sms() * 8 ? want : (int64_t)emo_num_sms() * 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (dl_dtype == EMO_BF16)
    ce_kernel<bf16><<<blocks, 256, 0, s>>>(logits, ld, tgt, rows, tgt_inner, tgt_stride_outer, tgt_stride_inner, V, ignore_index, count, gscale, loss_sum, ncorrect, pred, (bf16*)dlogits, ld_dl);
  else
    ce_kernel<float><<<blocks, 256, 0, s>>>(logits, ld, tgt, rows, tgt_inner, tgt_stride_outer, tgt_stride_inner, V, ignore_index, count, gscale, loss_sum, ncorrect, pred, (float*)dlogits, ld_dl);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

// ---------------------------------------------------------------------------------------------
// K12 grad-norm + Adam on flat buffers
// ---------------------------------------------------------------------------------------------
// Deterministic: block partials are written to a workspace and the LAST block to finish adds them up in block order,
// so every data-parallel replica gets bit-identical clip factors from bit-identical (all-reduced) gradients.  (A float
// atomicAdd per block gave last-bit differences between ranks: the replicas drifted apart by ulps per step.)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out,
                                                    float* __restrict__ partial, unsigned int* __restrict__ ticket) {
  float acc = 0.f;
  int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 << 2; i < n; ++i) acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float s[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    float t = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) t += __ldcg(partial + i);      // fixed assignment, fixed order
    t = warp_sum(t);
    if (threadIdx.x == 0) { *out += t; *ticket = 0u; }
  }
}
static int sumsq_workspace(float** partial, unsigned int** ticket, int blocks) {
  constexpr int MAX_DEV = 64, MAX_BLOCKS = 4096;
  static float* ws[MAX_DEV] = {nullptr};
  int dev = 0;
  EMO_CHECK_CUDA(cudaGetDevice(&dev));
  EMO_REQUIRE(dev < MAX_DEV && blocks <= MAX_BLOCKS, "emo_sumsq: device index / grid out of range");
  if (!ws[dev]) {
    EMO_CHECK_CUDA(cudaMalloc(&ws[dev], (MAX_BLOCKS + 1) * sizeof(float)));
    EMO_CHECK_CUDA(cudaMemset(ws[dev], 0, (MAX_BLOCKS + 1) * sizeof(float)));
  }
  *partial = ws[dev];
  *ticket = reinterpret_cast<unsigned int*>(ws[dev] + MAX_BLOCKS);
  return EMO_OK;
}
extern "C" int emo_sumsq(const float* g, int64_t n, float* out, void* stream) {
  if (n == 0) return EMO_OK;
  EMO_REQUIRE(((uintptr_t)g & 15) == 0, "emo_sumsq: buffer must be 16-byte aligned");
  int64_t want = (n / 4 + 255) / 256;
  int blocks = (int)(want < (int64_t)emo_num_sms() * 8 ? (want > 0 ? want : 1) : (int64_t)emo_num_sms() * 8);
  float* partial;
  unsigned int* ticket;
  int rc = sumsq_workspace(&partial, &ticket, blocks);      // one workspace per device: calls on one device must be stream-ordered
  if (rc) return rc;
  sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out, partial, ticket);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, bf16* __restrict__ pb, int64_t n, float lr,
                                                   float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                                   const float* __restrict__ gnorm_sq, float max_norm,
                                                   float grad_scale, int zero_grad) {
  float coef = grad_scale;
  if (max_norm > 0.f && gnorm_sq) {
    float nrm = sqrtf(gnorm_sq[0]) * grad_scale;
    float c = max_norm / (nrm + 1e-6f);
    coef *= (c < 1.f ? c : 1.f);
  }
  float step_size = lr / bc1;
  int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = ga[j] * coef;
      ma[j] = b1 * ma[j] + (1.f - b1) * gg;
      va[j] = b2 * va[j] + (1.f - b2) * gg * gg;
      float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0, 0, 0, 0);
    if (pb) {
      uint2 o;
      o.x = pack_bf16x2(pa[0], pa[1]);
      o.y = pack_bf16x2(pa[2], pa[3]);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
}
extern "C" int emo_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                             float beta1, float beta2, float eps, int64_t step, const float* gnorm_sq,
                             float max_norm, float grad_scale, int zero_grad, void* stream) {
  EMO_REQUIRE(n % 4 == 0, "emo_adam_step: n must be a multiple of 4 (pad the flat buffer)");
  EMO_REQUIRE(step >= 1, "emo_adam_step: step is 1-based");
  if (n == 0) return EMO_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t want = (n / 4 + 255) / 256;
  int blocks = (int)(want < (int64_t)emo_num_sms() * 8 ? want : (int64_t)emo_num_sms() * 8);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_bf16, n, lr, beta1, beta2, eps, (float)bc1,
                                                        (float)sqrt(bc2), gnorm_sq, max_norm, grad_scale, zero_grad);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = from_f<TD>(to_f(s[i]));
}
extern "C" int emo_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  if (n == 0) return EMO_OK;
  int blocks = (int)((n + 255) / 256 < (int64_t)emo_num_sms() * 16 ? (n + 255) / 256 : (int64_t)emo_num_sms() * 16);
  cudaStream_t s = (cudaStream_t)stream;
  if (src_dtype == EMO_F32 && dst_dtype == EMO_BF16) cast_kernel<float, bf16><<<blocks, 256, 0, s>>>((const float*)src, (bf16*)dst, n);
  else if (src_dtype == EMO_BF16 && dst_dtype == EMO_F32) cast_kernel<bf16, float><<<blocks, 256, 0, s>>>((const bf16*)src, (float*)dst, n);
  else if (src_dtype == EMO_F32 && dst_dtype == EMO_F32) cast_kernel<float, float><<<blocks, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else cast_kernel<bf16, bf16><<<blocks, 256, 0, s>>>((const bf16*)src, (bf16*)dst, n);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
