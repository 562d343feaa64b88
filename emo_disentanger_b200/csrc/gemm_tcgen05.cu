// K3/K6/K7/K10: bf16 dense projections on the 5th-gen tensor cores (sm_100a only).
//
// Persistent warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   cp.async.bulk.tensor (128B swizzle) -> 4/6-stage smem ring
//   warp 1      MMA issuer     tcgen05.mma.cta_group::1.kind::f16 (one elected lane), fp32
//                              accumulators in TMEM, double-buffered (2 x BN columns)
//   warps 2..9  epilogue       tcgen05.ld (32x32b.x32) -> bias / activation / dropout / residual
//                              -> 16-byte global stores, or fp32 atomic accumulate (split-K wgrad)
// Tile 128 x BN (BN = 256 or 128) x 64.  Three contractions share the kernel through the operand
// "major-ness" encoded in the TMA boxes, the smem descriptors and the instruction descriptor:
//   NT (forward)  A K-major,  B K-major          NN (dgrad)  A K-major,  B MN-major
//   TN (wgrad)    A MN-major, B MN-major (reduction over tokens, split-K + atomics)
#include "gemm_epilogue.cuh"
#include <cuda.h>

int emo_gemm_simt(int op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
                  void* C, int64_t ldc, int in_dtype, int out_dtype, const EpiParams& ep, cudaStream_t s);

namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int NUM_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long start = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - start > 8000000000LL) __trap();   // watchdog: never hang the GPU on a protocol bug
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (sm_100 UMMA): 128B swizzle, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct Params {
  int64_t M, N, K;
  void* C;
  int64_t ldc;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  EpiParams ep;
};


// ---- epilogue of one 32-column chunk of one accumulator row (lane = row) -------------------------
// Fast path: every branch below is warp-uniform and sits OUTSIDE the per-element loops; all global
// accesses are 16-byte vectors.  Order of operations = include/emo_b200.h (emo_epilogue).
template <typename TOut>
__device__ __forceinline__ void epi_chunk_vec(const uint32_t (&r)[32], int64_t m, int64_t nb, const Params& p) {
  const EpiParams& ep = p.ep;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (ep.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
  }
  if (ep.rowscale) {
    const float rs = ep.rowscale[m];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rs;
  }
  if (ep.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + nb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(b4 + j);
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (ep.act == EMO_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (ep.act == EMO_ACT_GELU_NEW) {
    if (ep.aux_out) {
      TOut* arow = reinterpret_cast<TOut*>(ep.aux_out) + m * ep.ld_aux + nb;
#pragma unroll
      for (int j = 0; j < 32; j += Vec<TOut>::N) {
        Vec<TOut> t;
#pragma unroll
        for (int i = 0; i < Vec<TOut>::N; ++i) t.v[i] = v[j + i];
        t.store(arow + j);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_new_f(v[j]);
  } else if (ep.act == EMO_ACT_RELU_MASK_BWD || ep.act == EMO_ACT_GELU_NEW_BWD) {
    const bf16* arow = reinterpret_cast<const bf16*>(ep.aux) + m * ep.ld_aux + nb;
    float a[32];
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      Vec<bf16> t;
      t.load(arow + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[j + i] = t.v[i];
    }
    if (ep.act == EMO_ACT_RELU_MASK_BWD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (a[j] != 0.f) ? v[j] * ep.aux_scale : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_new_grad_f(a[j]);
    }
  }
  if (ep.drop_thr) {
    const uint64_t e0 = (uint64_t)(m * ep.n_total + nb);
    if ((e0 & 1) == 0 && ((e0 >> 33) == ((e0 + 31) >> 33))) {
      // same value as emo_drop_hash(seed, e0 + j): the high half of the pair index is constant here
      const uint32_t base = (uint32_t)ep.seed ^ emo_mix32((uint32_t)(e0 >> 33) + (uint32_t)(ep.seed >> 32) + 0x7f4a7c15u);
      const uint32_t lo0 = (uint32_t)(e0 >> 1);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        uint32_t h = emo_mix32(((lo0 + (j >> 1)) * 0x9E3779B9u) ^ base);
        v[j] = ((h & 0xffffu) >= ep.drop_thr) ? v[j] * ep.keep_scale : 0.f;
        v[j + 1] = ((h >> 16) >= ep.drop_thr) ? v[j + 1] * ep.keep_scale : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = emo_drop_keep(ep.seed, e0 + j, ep.drop_thr) ? v[j] * ep.keep_scale : 0.f;
    }
  }
  if (ep.accumulate) {
    float* crow = reinterpret_cast<float*>(p.C) + m * p.ldc + nb;
#pragma unroll
    for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4*>(crow + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    return;
  }
  if (ep.residual) {
    const TOut* rrow = reinterpret_cast<const TOut*>(ep.residual) + m * ep.ld_res + nb;
#pragma unroll
    for (int j = 0; j < 32; j += Vec<TOut>::N) {
      Vec<TOut> t;
      t.load(rrow + j);
#pragma unroll
      for (int i = 0; i < Vec<TOut>::N; ++i) v[j + i] += t.v[i];
    }
  }
  TOut* crow = reinterpret_cast<TOut*>(p.C) + m * p.ldc + nb;
#pragma unroll
  for (int j = 0; j < 32; j += Vec<TOut>::N) {
    Vec<TOut> t;
#pragma unroll
    for (int i = 0; i < Vec<TOut>::N; ++i) t.v[i] = v[j + i];
    t.store(crow + j);
  }
}

// Slow path (ragged last columns, unaligned leading dims): per-element, fully guarded.
template <typename TOut>
__device__ __noinline__ void epi_chunk_generic(const uint32_t (&r)[32], int64_t m, int64_t nb, const Params& p) {
  const EpiParams& ep = p.ep;
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    if (nb + j >= p.N) break;
    float x = epi_full<bf16, TOut>(__uint_as_float(r[j]), m, nb + j, ep);
    if (ep.accumulate) atomicAdd(reinterpret_cast<float*>(p.C) + m * p.ldc + nb + j, x);
    else reinterpret_cast<TOut*>(p.C)[m * p.ldc + nb + j] = from_f<TOut>(x);
  }
}

template <int BN, bool A_MN, bool B_MN, typename TOut>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment required by the 128B swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.m_tiles * p.n_tiles;
  const int items = tiles * p.splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        int split = it / tiles, rem = it % tiles;
        int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * BN;
        int kb0 = split * p.kb_per_split;
        int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          int k0 = kb * BK;
          if (!A_MN) tma_load_2d(&tmA, full_bar(stage), sa, k0, m0);
          else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(&tmA, full_bar(stage), sa + i * 8192, m0 + 64 * i, k0);
          }
          if (!B_MN) tma_load_2d(&tmB, full_bar(stage), sb, k0, n0);
          else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(&tmB, full_bar(stage), sb + i * 8192, n0 + 64 * i, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        int split = it / tiles;
        int kb0 = split * p.kb_per_split;
        int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t ad = A_MN ? make_desc(sa + k * 2048, 8192, 1024) : make_desc(sa + k * 32, 0, 1024);
            uint64_t bd = B_MN ? make_desc(sb + k * 2048, 8192, 1024) : make_desc(sb + k * 32, 0, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---- epilogue warps 2..9: TMEM lane quadrant = warp % 4; the two warps of a quadrant split the
    //      tile's columns (first / second half) ----
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    const EpiParams& ep = p.ep;
    constexpr int VEC = 16 / sizeof(TOut);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec_ok = ((p.ldc % VEC) == 0) && al16(p.C) &&
                        (ep.residual == nullptr || (((ep.ld_res % VEC) == 0) && al16(ep.residual))) &&
                        (ep.bias == nullptr || al16(ep.bias)) &&
                        (ep.aux == nullptr || (((ep.ld_aux % 8) == 0) && al16(ep.aux))) &&
                        (ep.aux_out == nullptr || (((ep.ld_aux % VEC) == 0) && al16(ep.aux_out)));
    constexpr int CH_PER_WARP = BN / 64;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      int rem = it % tiles;
      int64_t m0 = (int64_t)(rem / p.n_tiles) * BM, n0 = (int64_t)(rem % p.n_tiles) * BN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int64_t m = m0 + quad * 32 + lane;
      const bool row_ok = m < p.M;
#pragma unroll 1
      for (int c = 0; c < CH_PER_WARP; ++c) {
        const int ch = half * CH_PER_WARP + c;
        const int64_t nb = n0 + ch * 32;
        if (nb >= p.N) break;   // warp-uniform
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + ch * 32, r);
        if (!row_ok) continue;
        if (vec_ok && nb + 32 <= p.N) epi_chunk_vec<TOut>(r, m, nb, p);
        else epi_chunk_generic<TOut>(r, m, nb, p);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor: `inner` contiguous elements, `outer` rows of stride ld elements; box {64, box_outer}
static int make_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { emo_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return EMO_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { emo_set_error("cuTensorMapEncodeTiled failed: %d (inner=%lld outer=%lld ld=%lld)", (int)r, (long long)inner, (long long)outer, (long long)ld); return EMO_ERR_CUDA; }
  return EMO_OK;
}

template <int BN, bool A_MN, bool B_MN, typename TOut>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p, cudaStream_t s) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + 256;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TOut>;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int items = p.m_tiles * p.n_tiles * p.splits;
  int grid = items < emo_num_sms() ? items : emo_num_sms();
  kern<<<grid, NUM_THREADS, smem, s>>>(tmA, tmB, p);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

template <int BN, typename TOut>
static int launch_op(int op, const CUtensorMap& a, const CUtensorMap& b, const Params& p, cudaStream_t s) {
  switch (op) {
    case EMO_GEMM_NT: return launch<BN, false, false, TOut>(a, b, p, s);
    case EMO_GEMM_NN: return launch<BN, false, true, TOut>(a, b, p, s);
    case EMO_GEMM_TN: return launch<BN, true, true, TOut>(a, b, p, s);
  }
  emo_set_error("emo_gemm: bad op %d", op);
  return EMO_ERR_ARG;
}

}  // namespace tc

static int g_force_simt = 0;
extern "C" void emo_gemm_force_simt(int on) { g_force_simt = on; }

extern "C" int emo_gemm(int op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                        int64_t ldb, void* C, int64_t ldc, int in_dtype, int out_dtype, const emo_epilogue* epi,
                        void* stream) {
  EMO_REQUIRE(op >= 0 && op <= 2, "emo_gemm: bad op %d", op);
  EMO_REQUIRE(M >= 0 && N >= 0 && K >= 0, "emo_gemm: negative dims");
  EMO_REQUIRE(A && B && C, "emo_gemm: null operand");
  if (M == 0 || N == 0) return EMO_OK;
  EpiParams ep = make_epi(epi, N);
  EMO_REQUIRE(!ep.accumulate || out_dtype == EMO_F32, "emo_gemm: accumulate needs an fp32 output");
  EMO_REQUIRE(K > 0 || ep.accumulate, "emo_gemm: K == 0 only makes sense with accumulate");
  if (K == 0) return EMO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  bool tc_ok = (in_dtype == EMO_BF16) && !g_force_simt && (lda % 8 == 0) && (ldb % 8 == 0) &&
               ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  if (!tc_ok) return emo_gemm_simt(op, M, N, K, A, lda, B, ldb, C, ldc, in_dtype, out_dtype, ep, s);

  using namespace tc;
  const int BN = (N % 256 == 0 || N >= 1024) ? 256 : 128;
  Params p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.ep = ep;
  p.m_tiles = (int)((M + BM - 1) / BM);
  p.n_tiles = (int)((N + BN - 1) / BN);
  p.kb_total = (int)((K + BK - 1) / BK);
  p.splits = 1;
  if (ep.accumulate) {
    int tiles = p.m_tiles * p.n_tiles;
    int want = (2 * emo_num_sms() + tiles - 1) / tiles;
    int maxs = p.kb_total / 8 > 0 ? p.kb_total / 8 : 1;
    p.splits = want < maxs ? want : maxs;
    if (p.splits < 1) p.splits = 1;
  }
  p.kb_per_split = (p.kb_total + p.splits - 1) / p.splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;

  CUtensorMap tmA, tmB;
  int rc;
  if (op == EMO_GEMM_TN) rc = make_map(&tmA, A, M, K, lda, 64);   // A stored [K][M]: inner = M
  else rc = make_map(&tmA, A, K, M, lda, BM);                      // A stored [M][K]: inner = K
  if (rc) return rc;
  if (op == EMO_GEMM_NT) rc = make_map(&tmB, B, K, N, ldb, BN);   // B stored [N][K]
  else rc = make_map(&tmB, B, N, K, ldb, 64);                      // B stored [K][N]: inner = N
  if (rc) return rc;

  if (out_dtype == EMO_BF16) return BN == 256 ? launch_op<256, bf16>(op, tmA, tmB, p, s) : launch_op<128, bf16>(op, tmA, tmB, p, s);
  if (out_dtype == EMO_F32) return BN == 256 ? launch_op<256, float>(op, tmA, tmB, p, s) : launch_op<128, float>(op, tmA, tmB, p, s);
  emo_set_error("emo_gemm: bad out dtype %d", out_dtype);
  return EMO_ERR_ARG;
}
