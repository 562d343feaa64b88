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
