// sm_100a PTX wrappers shared by the tcgen05 kernels outside the GEMM (FAVOR+ forward / backward, attention):
// mbarrier, TMA (cp.async.bulk.tensor), TMEM allocation / load / store, tcgen05.mma with the A operand in shared
// memory (SS) or in tensor memory (TS), shared-memory matrix descriptors.  Same encodings as gemm_tcgen05.cu, which
// the GEMM parity tests pin on hardware.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace tcp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive that releases shared memory this thread has READ (a "buffer free" signal to a TMA producer).  An mbarrier
// arrive does not wait for earlier ld.shared instructions whose results it does not use: MEASURED on B200, a plain
// arrive right after the loads let the next chunk's TMA overwrite rows that were still being read (a few rows per
// ~1000 CTAs got a wrong |q|^2).  `dep` must be computed from every loaded value; the barrier address is made to
// depend on it (the compared pattern is a NaN no arithmetic produces, so the offset is always 0).
__device__ __forceinline__ void mbar_arrive_after_reads(uint32_t bar, float dep) {
  const uint32_t off = (__float_as_uint(dep) == 0xFFFFFFFFu) ? 8u : 0u;
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + off) : "memory");
}
// parity wait; try_wait suspends in hardware up to the hint, the spin counter is a watchdog (a protocol bug traps
// after a few seconds instead of hanging the GPU)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
    if (++spins > (1u << 21)) __trap();
  }
}
// the same wait without a suspend-time hint: try_wait returns after the default (short) time limit and the loop
// polls again -- for waits on a tile-to-tile critical path, where waking from a long suspend costs more than polling
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int ID> __device__ __forceinline__ void named_bar_sync(int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}

// ---- TMA -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) { asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory"); }
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// global[tile] += shared[tile] (element type and add come from the tensor map: fp32); rows outside the tensor are clipped
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive columns: thread l of the warp gets TMEM lane (quadrant base + l), columns c .. c + 31
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- MMA -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: A is K-major, TMEM lane = row, 32-bit column c = elements (2c, 2c + 1) of the row
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// shared-memory matrix descriptor: 128B swizzle, descriptor version 1 (sm_100)
//   K-major  operand (rows = M or N, 128-byte rows of 64 bf16 along K): lbo unused, sbo = 1024 (8 rows);
//            K step of 16 elements = +32 bytes inside the swizzle atom
//   MN-major operand (rows = K, 128-byte rows of 64 bf16 along M or N): lbo = byte stride between 64-element MN
//            blocks, sbo = 1024 (8 K rows); K step of 16 rows = +2048 bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// instruction descriptor, kind::f16: bf16 x bf16 -> fp32, M = 128
// a_bf16 / b_bf16: operand formats (1 = bf16, 0 = fp16; the two may differ: the descriptor carries one field each)
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn, int m = 128, bool a_bf16 = true, bool b_bf16 = true) {
  return (1u << 4) | ((a_bf16 ? 1u : 0u) << 7) | ((b_bf16 ? 1u : 0u) << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// ---- warp-uniform issue path ------------------------------------------------------------------------------------
// The control warp runs its loop with ALL lanes (waits included) and only the instruction itself is given to one
// elected lane.  MEASURED on B200: with the whole loop under `if (lane == 0)` every descriptor lives in vector
// registers of a divergent region and each tcgen05.mma costs ~17 SASS instructions (R2UR moves plus an ELECT /
// BRA.U.ANY loop), ~72 clocks per MMA -- more than an N = 64, K = 16 MMA takes on the tensor pipe (32 clocks).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// descriptor halves: adding (bytes >> 4) to `lo` moves the start address (no carry out of the 14-bit field for any
// shared-memory address)
struct DescLH { uint32_t lo, hi; };
__device__ __forceinline__ DescLH make_desc_lh(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  DescLH d;
  d.lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
  d.hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
  return d;
}
__device__ __forceinline__ uint64_t desc_at(const DescLH& d, uint32_t byte_off) {
  return ((uint64_t)d.hi << 32) | (uint64_t)(d.lo + (byte_off >> 4));
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Omega' = Omega * 64^(-1/4) * log2(e) -> 16-bit [e][f] tile (128B swizzle).  fp16 (|Omega'| < 4, 11-bit significand)
// keeps the feature-map exponents 8x closer to the fp32 oracle than bf16 does; x stays bf16 (mixed-format MMA).
__device__ __forceinline__ void stage_omega(const float* omega, uint32_t tile, int tid, int nthreads, bool f16) {
  const float sc = 0.35355339059327373f * 1.4426950408889634f;
  for (int i = tid; i < 64 * 8; i += nthreads) {
    const int e = i >> 3, c = i & 7;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(omega + e * 64 + c * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(omega + e * 64 + c * 8 + 4));
    uint4 t;
    if (f16) {
      t.x = pack_f16x2(w0.x * sc, w0.y * sc); t.y = pack_f16x2(w0.z * sc, w0.w * sc);
      t.z = pack_f16x2(w1.x * sc, w1.y * sc); t.w = pack_f16x2(w1.z * sc, w1.w * sc);
    } else {
      t.x = pack_bf16x2(w0.x * sc, w0.y * sc); t.y = pack_bf16x2(w0.z * sc, w0.w * sc);
      t.z = pack_bf16x2(w1.x * sc, w1.y * sc); t.w = pack_bf16x2(w1.z * sc, w1.w * sc);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + (uint32_t)(e * 128 + ((c ^ (e & 7)) << 4))), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
  }
}
// MEASURED on B200: a kind::f16 MMA with a bf16 A operand and an fp16 B operand raises "illegal instruction" -- the two
// format fields of the instruction descriptor must agree.  The switch stays off; it documents the experiment.
static inline int favor_omega_f16() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("EMO_FAVOR_OMEGA_F16"); on = e ? atoi(e) : 0; }
  return on;
}
// byte offset of 16-byte chunk `c` of row `r` in a [rows][128 B] tile with the 128B swizzle (TMA / UMMA layout)
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 t;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr) : "memory");
  return t;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& t) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t t;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(addr) : "memory");
  return t;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- host: tensor maps -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// bf16 [B][T][row of `inner` elements] with token stride ld: box = {64, box_rows, 1}, 128B swizzle; rows past T read
// as zeros and are clipped on store (so a ragged last chunk needs no masking of the copies)
static inline int make_map_bt(CUtensorMap* map, const void* ptr, int64_t inner, int64_t T, int64_t B, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { emo_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return EMO_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    emo_set_error("cuTensorMapEncodeTiled (3d) failed: %d (inner=%lld T=%lld B=%lld ld=%lld)", (int)r, (long long)inner, (long long)T,
                  (long long)B, (long long)ld);
    return EMO_ERR_CUDA;
  }
  return EMO_OK;
}

// fp32 [B][T][row of `inner` elements], dense rows: box = {32 floats = 128 B, box_rows, 1}, 128B swizzle (the dQ
// accumulation target of the attention backward: cp.reduce.async.bulk.tensor adds tiles into it)
static inline int make_map_f32_bt(CUtensorMap* map, const void* ptr, int64_t inner, int64_t T, int64_t B, int box_rows, int box_cols = 32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { emo_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return EMO_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)T * (cuuint64_t)inner * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};       // 32 floats: 128B swizzle; 16 floats: 64B swizzle
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    emo_set_error("cuTensorMapEncodeTiled (3d, fp32) failed: %d (inner=%lld T=%lld B=%lld)", (int)r, (long long)inner, (long long)T, (long long)B);
    return EMO_ERR_CUDA;
  }
  return EMO_OK;
}

}  // namespace tcp
