// Block-level small-GEMM helper over shared-memory operands, used by the attention-type kernels
// (FAVOR+ chunked scan, causal softmax attention, rel-pos attention).
//
//   C[M x N] += A[M x K] . B[K x N]         BG_WARPS warps (8 by default, 16 on request) cooperate on one product.
//
// Storage flags describe how an operand sits in shared memory:
//   A_KMAJ = true : A stored [M][K] (K contiguous);  false: stored [K][M] (M contiguous)
//   B_KMAJ = true : B stored [N][K] (K contiguous);  false: stored [K][N] (N contiguous)
//
// Two implementations behind one interface:
//   T = bf16  : warp-level tensor-core MMA (mma.sync.m16n8k16, fp32 accumulate) with ldmatrix
//               (+ .trans for the non-K-major storages); rows padded by 8 elements (16 B).
//   T = float : SIMT fp32 FMA (the 1e-3 parity mode); rows padded by 1 element.
// Accumulators live in registers; `foreach` visits (row, col, value&) so callers never depend on
// the fragment layout.
#pragma once
#include "common.cuh"

#ifndef BG_WARPS
#define BG_WARPS 8       // warps per CTA cooperating on one product (a translation unit may set 16 before including)
#endif
constexpr int BG_THREADS = 32 * BG_WARPS;

template <typename T> struct BGPad;
template <> struct BGPad<bf16> { static constexpr int PAD = 8; };
template <> struct BGPad<float> { static constexpr int PAD = 1; };
template <typename T> __host__ __device__ constexpr int bg_ld(int cols) { return cols + BGPad<T>::PAD; }

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int M, int N, typename T> struct BlockGemm;

// ---------------------------------------------------------------------------------------------
// bf16 tensor-core implementation
// ---------------------------------------------------------------------------------------------
template <int M, int N> struct BlockGemm<M, N, bf16> {
  static_assert(M % 16 == 0 && N % 8 == 0, "tile shape");
  static constexpr int WM = (M >= 128) ? 8 : (M >= 64 ? 4 : ((M >= 32 || BG_WARPS > 8) ? 2 : 1));
  static constexpr int WN = BG_WARPS / WM;
  static_assert(M % (16 * WM) == 0 && N % (8 * WN) == 0, "warp split");
  static constexpr int MT = M / WM / 16;
  static constexpr int NT = N / WN / 8;
  float acc[MT][NT][4];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  }

  template <bool A_KMAJ, bool B_KMAJ>
  __device__ __forceinline__ void mma(const bf16* A, int lda, const bf16* B, int ldb, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_base = (warp / WN) * (MT * 16);
    const int n_base = (warp % WN) * (NT * 8);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int m0 = m_base + mt * 16;
        if (A_KMAJ) {
          int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
          int kk = k0 + (lane >> 4) * 8;
          ldsm_x4(a[mt][0], a[mt][1], a[mt][2], a[mt][3], A + row * lda + kk);
        } else {
          int mat = lane >> 3;
          int kk = k0 + (lane & 7) + (mat >> 1) * 8;
          int mm = m0 + (mat & 1) * 8;
          ldsm_x4_t(a[mt][0], a[mt][1], a[mt][2], a[mt][3], A + kk * lda + mm);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {
        uint32_t b[4];
        int n0 = n_base + nt * 8;
        if (nt + 1 < NT) {
          if (B_KMAJ) {
            int nn = n0 + (lane & 7) + (lane >> 4) * 8;
            int kk = k0 + ((lane >> 3) & 1) * 8;
            ldsm_x4(b[0], b[1], b[2], b[3], B + nn * ldb + kk);
          } else {
            int kk = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
            int nn = n0 + (lane >> 4) * 8;
            ldsm_x4_t(b[0], b[1], b[2], b[3], B + kk * ldb + nn);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16_16816(acc[mt][nt], a[mt], b[0], b[1]);
            mma_bf16_16816(acc[mt][nt + 1], a[mt], b[2], b[3]);
          }
        } else {
          int l = lane & 15;
          if (B_KMAJ) {
            int nn = n0 + (l & 7);
            int kk = k0 + ((l >> 3) & 1) * 8;
            ldsm_x2(b[0], b[1], B + nn * ldb + kk);
          } else {
            int kk = k0 + (l & 7) + ((l >> 3) & 1) * 8;
            ldsm_x2_t(b[0], b[1], B + kk * ldb + n0);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) mma_bf16_16816(acc[mt][nt], a[mt], b[0], b[1]);
        }
      }
    }
  }

  // compile-time K: the k loop is fully unrolled, so ptxas can hoist the ldmatrix of step k+1 above the MMAs of
  // step k (with a run-time trip count every step is load -> wait -> mma)
  template <bool A_KMAJ, bool B_KMAJ, int K>
  __device__ __forceinline__ void mma_k(const bf16* A, int lda, const bf16* B, int ldb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_base = (warp / WN) * (MT * 16);
    const int n_base = (warp % WN) * (NT * 8);
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int m0 = m_base + mt * 16;
        if (A_KMAJ) {
          int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
          int kk = k0 + (lane >> 4) * 8;
          ldsm_x4(a[mt][0], a[mt][1], a[mt][2], a[mt][3], A + row * lda + kk);
        } else {
          int mat = lane >> 3;
          int kk = k0 + (lane & 7) + (mat >> 1) * 8;
          int mm = m0 + (mat & 1) * 8;
          ldsm_x4_t(a[mt][0], a[mt][1], a[mt][2], a[mt][3], A + kk * lda + mm);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {
        uint32_t b[4];
        int n0 = n_base + nt * 8;
        if (nt + 1 < NT) {
          if (B_KMAJ) {
            int nn = n0 + (lane & 7) + (lane >> 4) * 8;
            int kk = k0 + ((lane >> 3) & 1) * 8;
            ldsm_x4(b[0], b[1], b[2], b[3], B + nn * ldb + kk);
          } else {
            int kk = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
            int nn = n0 + (lane >> 4) * 8;
            ldsm_x4_t(b[0], b[1], b[2], b[3], B + kk * ldb + nn);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16_16816(acc[mt][nt], a[mt], b[0], b[1]);
            mma_bf16_16816(acc[mt][nt + 1], a[mt], b[2], b[3]);
          }
        } else {
          int l = lane & 15;
          if (B_KMAJ) {
            int nn = n0 + (l & 7);
            int kk = k0 + ((l >> 3) & 1) * 8;
            ldsm_x2(b[0], b[1], B + nn * ldb + kk);
          } else {
            int kk = k0 + (l & 7) + ((l >> 3) & 1) * 8;
            ldsm_x2_t(b[0], b[1], B + kk * ldb + n0);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) mma_bf16_16816(acc[mt][nt], a[mt], b[0], b[1]);
        }
      }
    }
  }

  template <typename F> __device__ __forceinline__ void foreach (F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_base = (warp / WN) * (MT * 16);
    const int n_base = (warp % WN) * (NT * 8);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int row = m_base + mt * 16 + (lane >> 2) + ((i >> 1) ? 8 : 0);
          int col = n_base + nt * 8 + (lane & 3) * 2 + (i & 1);
          f(row, col, acc[mt][nt][i]);
        }
  }
  // pairs of adjacent columns (col even): lets callers emit one packed bf16x2 shared-memory store per pair
  template <typename F> __device__ __forceinline__ void foreach2(F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_base = (warp / WN) * (MT * 16);
    const int n_base = (warp % WN) * (NT * 8);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
          int row = m_base + mt * 16 + (lane >> 2) + (i ? 8 : 0);
          int col = n_base + nt * 8 + (lane & 3) * 2;
          f(row, col, acc[mt][nt][i], acc[mt][nt][i + 1]);
        }
  }
};

// ---------------------------------------------------------------------------------------------
// fp32 SIMT implementation: 16 x 16 thread grid, thread (ty, tx) owns rows ty+16i, cols tx+16j
// ---------------------------------------------------------------------------------------------
template <int M, int N> struct BlockGemm<M, N, float> {
  static_assert(M % 16 == 0 && N % 16 == 0, "tile shape");
  static_assert(BG_WARPS == 8 || M < 0, "the fp32 SIMT implementation is a 16 x 16 thread grid");
  static constexpr int MT = M / 16;
  static constexpr int NT = N / 16;
  float acc[MT][NT];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  }

  template <bool A_KMAJ, bool B_KMAJ>
  __device__ __forceinline__ void mma(const float* A, int lda, const float* B, int ldb, int K) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = A_KMAJ ? A[(ty + 16 * i) * lda + k] : A[k * lda + ty + 16 * i];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = B_KMAJ ? B[(tx + 16 * j) * ldb + k] : B[k * ldb + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  template <bool A_KMAJ, bool B_KMAJ, int K>
  __device__ __forceinline__ void mma_k(const float* A, int lda, const float* B, int ldb) { mma<A_KMAJ, B_KMAJ>(A, lda, B, ldb, K); }

  template <typename F> __device__ __forceinline__ void foreach (F f) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) f(ty + 16 * i, tx + 16 * j, acc[i][j]);
  }
};

// store two adjacent elements (col even) of a shared-memory tile
__device__ __forceinline__ void st_pair(bf16* p, float x0, float x1) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(x0, x1); }
__device__ __forceinline__ void st_pair(float* p, float x0, float x1) { p[0] = x0; p[1] = x1; }
