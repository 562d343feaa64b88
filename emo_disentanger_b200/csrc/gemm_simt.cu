// SIMT fp32-accumulate GEMM: the 1e-3 parity mode (fp32 operands) and the small/odd-shape path for
// bf16 operands that the tcgen05 kernel does not take (unaligned leading dims).  64x64x16 tiles,
// 256 threads, 4x4 outputs per thread, all three contractions (NT / NN / TN), full epilogue.
#include "gemm_epilogue.cuh"

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename TIn, typename TOut, int OP>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int64_t M, int64_t N, int64_t K, const TIn* __restrict__ A,
                                                        int64_t lda, const TIn* __restrict__ B, int64_t ldb,
                                                        TOut* __restrict__ C, int64_t ldc, EpiParams ep,
                                                        int64_t k_per_split) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += SG_BK) {
    // A tile: logical A(m, k).  NT/NN: stored [M][K]; TN: stored [K][M]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = threadIdx.x + i * 256;
      int mm, kk;
      if (OP == EMO_GEMM_TN) { kk = idx / SG_BM; mm = idx % SG_BM; }
      else { mm = idx / SG_BK; kk = idx % SG_BK; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < kend) v = to_f(OP == EMO_GEMM_TN ? A[gk * lda + gm] : A[gm * lda + gk]);
      As[kk][mm] = v;
    }
    // B tile: logical B(k, n).  NT: stored [N][K]; NN/TN: stored [K][N]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = threadIdx.x + i * 256;
      int nn, kk;
      if (OP == EMO_GEMM_NT) { nn = idx / SG_BK; kk = idx % SG_BK; }
      else { kk = idx / SG_BN; nn = idx % SG_BN; }
      int64_t gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < kend) v = to_f(OP == EMO_GEMM_NT ? B[gn * ldb + gk] : B[gk * ldb + gn]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < M && n < N) {
        float v = epi_full<TIn, TOut>(acc[i][j], m, n, ep);
        if (ep.accumulate) atomicAdd(reinterpret_cast<float*>(C) + m * ldc + n, v);
        else C[m * ldc + n] = from_f<TOut>(v);
      }
    }
}

template <typename TIn, typename TOut>
static int simt_dispatch(int op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                         int64_t ldb, void* C, int64_t ldc, const EpiParams& ep, cudaStream_t s) {
  dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM), 1);
  int64_t kps = K;
  if (ep.accumulate) {   // split-K over the token dimension for weight gradients
    int64_t tiles = (int64_t)grid.x * grid.y;
    int64_t want = (4LL * emo_num_sms() + tiles - 1) / tiles;
    int64_t maxs = (K + 255) / 256;
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    kps = ((K + want - 1) / want + SG_BK - 1) / SG_BK * SG_BK;
    grid.z = (unsigned)((K + kps - 1) / kps);
  }
  const TIn* a = (const TIn*)A; const TIn* b = (const TIn*)B; TOut* c = (TOut*)C;
  switch (op) {
    case EMO_GEMM_NT: gemm_simt_kernel<TIn, TOut, EMO_GEMM_NT><<<grid, 256, 0, s>>>(M, N, K, a, lda, b, ldb, c, ldc, ep, kps); break;
    case EMO_GEMM_NN: gemm_simt_kernel<TIn, TOut, EMO_GEMM_NN><<<grid, 256, 0, s>>>(M, N, K, a, lda, b, ldb, c, ldc, ep, kps); break;
    case EMO_GEMM_TN: gemm_simt_kernel<TIn, TOut, EMO_GEMM_TN><<<grid, 256, 0, s>>>(M, N, K, a, lda, b, ldb, c, ldc, ep, kps); break;
    default: emo_set_error("emo_gemm: bad op %d", op); return EMO_ERR_ARG;
  }
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

int emo_gemm_simt(int op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
                  void* C, int64_t ldc, int in_dtype, int out_dtype, const EpiParams& ep, cudaStream_t s) {
  if (in_dtype == EMO_F32 && out_dtype == EMO_F32) return simt_dispatch<float, float>(op, M, N, K, A, lda, B, ldb, C, ldc, ep, s);
  if (in_dtype == EMO_BF16 && out_dtype == EMO_BF16) return simt_dispatch<bf16, bf16>(op, M, N, K, A, lda, B, ldb, C, ldc, ep, s);
  if (in_dtype == EMO_BF16 && out_dtype == EMO_F32) return simt_dispatch<bf16, float>(op, M, N, K, A, lda, B, ldb, C, ldc, ep, s);
  emo_set_error("emo_gemm: unsupported dtype combination in=%d out=%d", in_dtype, out_dtype);
  return EMO_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// Skinny NT GEMM for the decode step (M <= 8 rows = sequences in flight): C[M,N] = A[M,K] . B[N,K]^T.
// Weight-streaming bound (every weight byte is read once per token: 2*N*K bytes), so the kernel is a
// GEMV: one warp per output column n streams the K-contiguous weight row with 16-byte loads (all of a
// lane's chunks in flight at once), the M activation rows sit in shared memory, fp32 accumulate,
// butterfly reduce, full epilogue on lane 0..M-1.  N/4 CTAs of 4 warps spread the rows over all SMs.
// ---------------------------------------------------------------------------------------------
constexpr int SK_MAXM = 8;
constexpr int SK_WARPS = 4;

template <typename TOut>
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_nt_kernel(int M, int64_t N, int K, const bf16* __restrict__ A,
                                                                       int64_t lda, const bf16* __restrict__ B, int64_t ldb,
                                                                       TOut* __restrict__ C, int64_t ldc, EpiParams ep) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  bf16* As = reinterpret_cast<bf16*>(sk_smem);           // [M][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kv = K >> 3;                                  // 16-byte vectors per row
  // the weight row does not depend on the previous kernel's output: get its first vectors in flight before
  // anything else (the step is a chain of small dependent launches; every exposed latency counts)
  const int64_t n = (int64_t)blockIdx.x * SK_WARPS + warp;
  const bf16* brow = B + (n < N ? n : 0) * ldb;
  uint4 bw0[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int c = lane + 32 * u;
    bw0[u] = (c < kv) ? __ldg(reinterpret_cast<const uint4*>(brow) + c) : make_uint4(0, 0, 0, 0);
  }
  // constants of the epilogue / prologue ride with the weight prefetch: every dependent global round trip that
  // can be taken off the critical path of these ~4 us kernels counts
  const float bias_n = (ep.bias && n < N) ? __ldg(ep.bias + n) : 0.f;
  float4 lg[4], lb[4];
  if (ep.ln_gamma) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = (lane + 32 * h) * 8;
      lg[2 * h] = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + c0));
      lg[2 * h + 1] = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + c0 + 4));
      lb[2 * h] = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + c0));
      lb[2 * h + 1] = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + c0 + 4));
    }
  }
  pdl_trigger();
  pdl_wait();              // everything below reads what the previous kernel of the step wrote
  // residual element of (row = lane, column n): in flight together with the activation rows
  float res_mn = 0.f;
  if (ep.residual && lane < M && n < N) res_mn = to_f(reinterpret_cast<const TOut*>(ep.residual)[(int64_t)lane * ep.ld_res + n]);
  for (int i0 = threadIdx.x; i0 < M * kv; i0 += SK_WARPS * 32 * 4) {      // 4 activation vectors in flight per thread
    uint4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int i = i0 + u * SK_WARPS * 32;
      if (i < M * kv) t[u] = *reinterpret_cast<const uint4*>(A + (int64_t)(i / kv) * lda + (i % kv) * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int i = i0 + u * SK_WARPS * 32;
      if (i < M * kv) reinterpret_cast<uint4*>(As)[i] = t[u];
    }
  }
  __syncthreads();
  if (ep.ln_gamma) {
    // LayerNorm prologue (K == 512): the rows are tiny, so every CTA normalises its own copy in shared memory
    // (fp32 statistics, bf16 result -- the numbers emo_ln_fwd writes); CTA 0 also stores them for the residual
    // branch of the next projection.  Saves a launch per LayerNorm in the launch-latency-bound decode step.
    for (int m = warp; m < M; m += SK_WARPS) {
      uint4* row = reinterpret_cast<uint4*>(As + m * K);
      float x[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 t = row[lane + 32 * h];
        unpack_bf16x2(t.x, x[8 * h], x[8 * h + 1]); unpack_bf16x2(t.y, x[8 * h + 2], x[8 * h + 3]);
        unpack_bf16x2(t.z, x[8 * h + 4], x[8 * h + 5]); unpack_bf16x2(t.w, x[8 * h + 6], x[8 * h + 7]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) sum += x[j];
      const float mu = warp_sum(sum) * (1.f / 512.f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { float d = x[j] - mu; q += d * d; }
      const float rs = rsqrtf(warp_sum(q) * (1.f / 512.f) + 1e-5f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = (lane + 32 * h) * 8;
        float y[8];
        const float gq[8] = {lg[2 * h].x, lg[2 * h].y, lg[2 * h].z, lg[2 * h].w, lg[2 * h + 1].x, lg[2 * h + 1].y, lg[2 * h + 1].z, lg[2 * h + 1].w};
        const float bq[8] = {lb[2 * h].x, lb[2 * h].y, lb[2 * h].z, lb[2 * h].w, lb[2 * h + 1].x, lb[2 * h + 1].y, lb[2 * h + 1].z, lb[2 * h + 1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (x[8 * h + j] - mu) * rs * gq[j] + bq[j];
        uint4 t;
        t.x = pack_bf16x2(y[0], y[1]); t.y = pack_bf16x2(y[2], y[3]); t.z = pack_bf16x2(y[4], y[5]); t.w = pack_bf16x2(y[6], y[7]);
        row[lane + 32 * h] = t;
        if (blockIdx.x == 0 && ep.ln_out) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.ln_out) + (int64_t)m * ep.ld_ln + c0) = t;
      }
    }
    __syncthreads();
  }
  if (n >= N) return;
  float acc[SK_MAXM];
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) acc[m] = 0.f;
  for (int c0 = lane; c0 < kv; c0 += 32 * 4) {            // 4 weight vectors in flight per lane per trip
    uint4 bw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int c = c0 + 32 * u;
      if (c0 == lane) bw[u] = bw0[u];                      // first trip: prefetched above
      else bw[u] = (c < kv) ? __ldg(reinterpret_cast<const uint4*>(brow) + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int c = c0 + 32 * u;
      if (c >= kv) break;
      float b[8];
      unpack_bf16x2(bw[u].x, b[0], b[1]); unpack_bf16x2(bw[u].y, b[2], b[3]);
      unpack_bf16x2(bw[u].z, b[4], b[5]); unpack_bf16x2(bw[u].w, b[6], b[7]);
#pragma unroll
      for (int m = 0; m < SK_MAXM; ++m) {
        if (m < M) {
          uint4 aw = reinterpret_cast<const uint4*>(As)[m * kv + c];
          float a[8];
          unpack_bf16x2(aw.x, a[0], a[1]); unpack_bf16x2(aw.y, a[2], a[3]);
          unpack_bf16x2(aw.z, a[4], a[5]); unpack_bf16x2(aw.w, a[6], a[7]);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[m] = fmaf(a[j], b[j], acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) {
    if (m < M) acc[m] = warp_sum(acc[m]);
  }
  // lane m finishes row m (epilogue reads bias / residual / aux for one element)
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) {
    if (m < M && lane == m) {
      EpiParams e2 = ep;                 // bias and residual were fetched up front
      e2.bias = nullptr;
      e2.residual = nullptr;
      float v = acc[m] * ep.alpha;
      if (ep.rowscale) v *= ep.rowscale[m];
      e2.alpha = 1.f;
      e2.rowscale = nullptr;
      v = epi_full<bf16, TOut>(v + bias_n, m, n, e2) + res_mn;
      C[(int64_t)m * ldc + n] = from_f<TOut>(v);
    }
  }
}

int emo_gemm_skinny_nt(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                       int64_t ldc, int out_dtype, const EpiParams& ep, cudaStream_t s) {
  const size_t smem = (size_t)M * K * sizeof(bf16);
  const unsigned grid = (unsigned)((N + SK_WARPS - 1) / SK_WARPS);
  if (out_dtype == EMO_BF16)
    EMO_CHECK_CUDA(emo_launch_dep(gemm_skinny_nt_kernel<bf16>, dim3(grid), dim3(SK_WARPS * 32), smem, s, (int)M, N, (int)K,
                                  (const bf16*)A, lda, (const bf16*)B, ldb, (bf16*)C, ldc, ep));
  else
    EMO_CHECK_CUDA(emo_launch_dep(gemm_skinny_nt_kernel<float>, dim3(grid), dim3(SK_WARPS * 32), smem, s, (int)M, N, (int)K,
                                  (const bf16*)A, lda, (const bf16*)B, ldb, (float*)C, ldc, ep));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
