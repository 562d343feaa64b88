// FAVOR+ backward on the 5th-gen tensor cores (bf16, sm_100a): every product of the reverse pass is a tcgen05.mma with
// its accumulator in TMEM; q / k / v / out / dout chunks are staged by TMA (128B swizzle).  Same math, interface and
// workspace layout as favor_bwd2_kernel (mma.sync), which stays as the A/B switch EMO_FAVOR_TC_BWD=0.
//
// Backward of fast_transformers CausalLinearAttention + Favor (causal_dot_product's backward and the feature map's),
// stage2_accompaniment/model/fast_transformer_decoder.py:28-38.
//
// One CTA per SM walks one (batch, head, segment) in REVERSE in chunks of 128 tokens.  With G = dout / den,
// gd = -(dout . out) / den (the normaliser's gradient), P = phi(q) phi(k)^T and dP_ij = G_i . v_j + gd_i (both causal):
//
//   d phi(q) = dP phi(k) + G S_prev^T + gd z_prev^T          S_prev, z_prev: prefix state before the chunk
//   d phi(k) = dP^T phi(q) + V R^T + 1 rz^T                  R = sum_{later} phi(q)^T G, rz = sum_{later} gd phi(q)
//   dv       = P^T G + phi(k) R
//   dx       = (d phi * phi -> du, do) : du Om'^T ln2 + do s^2 x         for x = q and k
//   R += phi(q)^T G ; S_prev = S - phi(k)^T V (rolled back from the forward's final state; kept negated in TMEM)
//
// The rank-one terms (gd, z, rz: the "ones column" of the reference formulation) live on the CUDA cores as fp32
// vectors, so every MMA is N = 64 / 128, K = 64 / 128.  The transposed tiles (P^T, dP^T) are computed as their own
// products (phi(k) phi(q)^T, V G^T) so that all masked tiles are K-major A operands held in TENSOR MEMORY (packed
// bf16 written in place over their fp32 accumulators); d U_q / d U_k are TMEM A operands too.
// TMEM (512 columns): R [0,64) | -S [64,128) | W1 [128,256) | W0 [256,384) | W2 [384,512); each W holds a 128-column
// score tile, then its packed copy in the low 64 columns and a 64-column accumulator ("hole") in the high ones.
//
// Threads: 8 worker warps + 1 control warp (TMA + MMA issue).  A worker thread is (g, r): r = TMEM lane = row of
// every tile, g = which HALF of the tile's columns it handles -- both warp groups work on every tile of every phase
// (the kernel is bound by instruction issue on the element-wise passes, not by the tensor pipe), through one code
// path.  Four CTA-wide barriers per chunk separate the five MMA batches; inside a phase the workers consume the
// batch's results in the order the control warp commits them, so tensor-core time hides behind the previous tile.
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace favor3b {
using namespace tcp;

constexpr int C = 128, FE = 64, FM = 128, FV = 80, NT = 288;
constexpr float F_S2 = 0.125f, F_HALF_LOG_M = 2.4260151319598084f, K2 = 1.4426950408889634f, KINV = 0.6931471805599453f;
constexpr uint32_t TILE = 16384;
constexpr uint32_t OFF_XQ = 0, OFF_XK = TILE, OFF_XV = 2 * TILE /* x2 */, OFF_XO = 4 * TILE, OFF_XD = 5 * TILE, OFF_PQ = 6 * TILE /* x2 */,
                   OFF_PK = 8 * TILE /* x2 */, OFF_G = 10 * TILE, OFF_SB = 11 * TILE, OFF_RB = 12 * TILE, OFF_OM = 13 * TILE /* 8 KB */,
                   OFF_VEC = 13 * TILE + 8192;
// fp32 vectors: z[128] rz[2][128] gd[128] sp[2][2][128] part[8][128]
constexpr uint32_t V_Z = 0, V_RZ = 512, V_GD = 1536, V_SP = 2048, V_PART = 4096, VEC_BYTES = 4096 + 4096;
constexpr uint32_t OFF_BAR = OFF_VEC + VEC_BYTES, SMEM_USED = OFF_BAR + 256;
constexpr int SMEM_BYTES = SMEM_USED + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");
constexpr uint32_t T_R = 0, T_NS = 64, T_W1 = 128, T_W0 = 256, T_W2 = 384, T_COLS = 512;

// mbarriers (8 bytes each, from OFF_BAR)
enum { B_XQK = 0, B_OD, B_V0, B_V1, F_X, F_OD, F_PHI, M_U, M_C1, M_C2, M_C3, M_NS, M_E1, M_E2, M_DXQ, M_G, M_R, M_DXK, N_BARS };
// named barriers: 1 = whole CTA (phase boundaries), 2 = group 1 only, 3..5 = load/store hand-over of the three masked
// tiles (group 0 arrives, group 1 waits), 6 = all workers
struct Params {
  const bf16* q; const bf16* k; int64_t ld;
  const float* omega; const float* den;
  const float* seg_states; const float* seg_rstates;
  bf16* dq; bf16* dk; bf16* dv; int64_t ld_d;
  int nseg, seg_chunks, fwd_nseg, ratio, T, H, items;
  int debug;
};

template <int ID> __device__ __forceinline__ void named_bar_arrive(int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
// 32 fp32 accumulator values -> 4 swizzled 16-byte chunks (chunks c0 .. c0 + 3 of row `row`)
__device__ __forceinline__ void pack_row32(uint32_t tile, int row, int c0, const uint32_t (&r)[32], float scale) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint4 t;
    t.x = pack_bf16x2(__uint_as_float(r[8 * cc]) * scale, __uint_as_float(r[8 * cc + 1]) * scale);
    t.y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]) * scale, __uint_as_float(r[8 * cc + 3]) * scale);
    t.z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]) * scale, __uint_as_float(r[8 * cc + 5]) * scale);
    t.w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]) * scale, __uint_as_float(r[8 * cc + 7]) * scale);
    sts128(tile + sw128(row, c0 + cc), t);
  }
}
__device__ __forceinline__ float row_sumsq(uint32_t tile, int row) {
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 a = lds128(tile + sw128(row, c));
    float f0, f1;
    unpack_bf16x2(a.x, f0, f1); ss += f0 * f0 + f1 * f1; unpack_bf16x2(a.y, f0, f1); ss += f0 * f0 + f1 * f1;
    unpack_bf16x2(a.z, f0, f1); ss += f0 * f0 + f1 * f1; unpack_bf16x2(a.w, f0, f1); ss += f0 * f0 + f1 * f1;
  }
  return ss;
}
// phi of one row: U (64 fp32 columns at tmem address `tu`) -> bf16 [token][feature] tile (two 16 KB blocks: exp(+u - o) | exp(-u - o))
__device__ __forceinline__ void phi_row_to_smem(uint32_t tu, uint32_t tile, int row, float o) {
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld32_issue(tu + half * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      uint4 tp, tm;
      uint32_t* pp = reinterpret_cast<uint32_t*>(&tp);
      uint32_t* pm = reinterpret_cast<uint32_t*>(&tm);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float u0 = __uint_as_float(r[8 * cc + 2 * e]), u1 = __uint_as_float(r[8 * cc + 2 * e + 1]);
        pp[e] = pack_bf16x2(ex2(u0 - o), ex2(u1 - o));
        pm[e] = pack_bf16x2(ex2(-u0 - o), ex2(-u1 - o));
      }
      sts128(tile + sw128(row, half * 4 + cc), tp);
      sts128(tile + TILE + sw128(row, half * 4 + cc), tm);
    }
  }
}
// column sums over the 128 token rows of a [token][128 features] tile, optionally weighted per row; NTH = 128 or 256
// threads.  thread t -> 16-byte chunk column (t & 15) of the two blocks, 128 * 16 / NTH rows; partial sums to part[8][128]
template <int NTH>
__device__ __forceinline__ void colsum_partial(uint32_t tile, int t, const float* w, float* part) {
  constexpr int ROWS = 128 * 16 / NTH;
  const int cc = t & 15, blk = cc >> 3, c = cc & 7, rg = t >> 4;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
  for (int r = rg * ROWS; r < rg * ROWS + ROWS; ++r) {
    const uint4 a = lds128(tile + blk * TILE + sw128(r, c));
    const float wr = w ? w[r] : 1.f;
    float f0, f1;
    unpack_bf16x2(a.x, f0, f1); acc[0] += wr * f0; acc[1] += wr * f1;
    unpack_bf16x2(a.y, f0, f1); acc[2] += wr * f0; acc[3] += wr * f1;
    unpack_bf16x2(a.z, f0, f1); acc[4] += wr * f0; acc[5] += wr * f1;
    unpack_bf16x2(a.w, f0, f1); acc[6] += wr * f0; acc[7] += wr * f1;
  }
  int slot = rg;
  if (NTH == 256) {                        // lanes l and l ^ 16 hold the same chunk column, adjacent row groups
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    slot = rg >> 1;
    if (t & 16) return;
  }
  float4* dst = reinterpret_cast<float4*>(part + slot * 128 + blk * 64 + c * 8);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}
__device__ __forceinline__ float part_total(const float* part, int f) {
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) s += part[g * 128 + f];
  return s;
}

// One masked score tile of the chunk: this thread's two 32-column pieces (2 g, 2 g + 1) of row r -> packed bf16 written
// in place over the tile's low columns [32 g, 32 g + 32).  MODE 0: keep col >= row (P^T); MODE 1: + gd of the row, keep
// col <= row (dP); MODE 2: + gd of the column, keep col >= row (dP^T).  Group 1's packed words land on fp32 columns
// that belong to group 0's pieces, hence the arrive / wait hand-over between load and store.
template <int MODE, int BAR>
__device__ __forceinline__ void mask_tile(uint32_t tt, int g, int wq, int r, float gdr, const float* gd) {
  uint32_t pk[32];
#pragma unroll
  for (int pp = 0; pp < 2; ++pp) {
    const int pc = 2 * g + pp;
    const bool needed = (MODE == 1) ? (pc <= wq) : (pc >= wq);           // warp-uniform
    if (needed) {
      uint32_t v[32];
      tmem_ld32_issue(tt + pc * 32, v);
      tmem_ld_wait();
      const bool diag = pc == wq;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float a0 = __uint_as_float(v[2 * q]), a1 = __uint_as_float(v[2 * q + 1]);
        const int c0 = pc * 32 + 2 * q;
        if (MODE == 1) { a0 += gdr; a1 += gdr; }
        if (MODE == 2) { const float2 gq = *reinterpret_cast<const float2*>(gd + c0); a0 += gq.x; a1 += gq.y; }
        if (diag) {
          if (MODE == 1) { a0 = c0 <= r ? a0 : 0.f; a1 = c0 + 1 <= r ? a1 : 0.f; }
          else { a0 = c0 >= r ? a0 : 0.f; a1 = c0 + 1 >= r ? a1 : 0.f; }
        }
        pk[pp * 16 + q] = pack_bf16x2(a0, a1);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) pk[pp * 16 + q] = 0u;
    }
  }
  tc_fence_before();
  if (g == 0) named_bar_arrive<BAR>(256);
  else named_bar_sync<BAR>(256);
  tc_fence_after();
  tmem_st32(tt + g * 32, pk);
}

// d phi of this thread's row, features [32 g, 32 g + 32) of both signs (accumulators at tp + 32 g / tm + 32 g):
//   w = (d phi + coef * vec) * phi ;  du = w+ - w- -> 16 packed words ;  returns sum(w) (partial over the 64 features)
__device__ __forceinline__ float dphi_half(uint32_t tp, uint32_t tm, uint32_t phi_tile, int row, int g, const float* vec, float coef,
                                           uint32_t (&du)[16]) {
  float s0 = 0.f, s1 = 0.f;
  const float4* v4 = reinterpret_cast<const float4*>(vec);
  uint32_t rp[32], rm[32];
  tmem_ld32_issue(tp + g * 32, rp);
  tmem_ld32_issue(tm + g * 32, rm);
  tmem_ld_wait();
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const uint4 php = lds128(phi_tile + sw128(row, g * 4 + cc)), phm = lds128(phi_tile + TILE + sw128(row, g * 4 + cc));
    const float4 vpa = v4[g * 8 + 2 * cc], vpb = v4[g * 8 + 2 * cc + 1];
    const float4 vma = v4[16 + g * 8 + 2 * cc], vmb = v4[16 + g * 8 + 2 * cc + 1];
    const uint32_t pw[4] = {php.x, php.y, php.z, php.w}, mw[4] = {phm.x, phm.y, phm.z, phm.w};
    const float vp[8] = {vpa.x, vpa.y, vpa.z, vpa.w, vpb.x, vpb.y, vpb.z, vpb.w};
    const float vm[8] = {vma.x, vma.y, vma.z, vma.w, vmb.x, vmb.y, vmb.z, vmb.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float p0, p1, m0, m1;
      unpack_bf16x2(pw[e], p0, p1);
      unpack_bf16x2(mw[e], m0, m1);
      const float wp0 = (__uint_as_float(rp[8 * cc + 2 * e]) + coef * vp[2 * e]) * p0;
      const float wp1 = (__uint_as_float(rp[8 * cc + 2 * e + 1]) + coef * vp[2 * e + 1]) * p1;
      const float wm0 = (__uint_as_float(rm[8 * cc + 2 * e]) + coef * vm[2 * e]) * m0;
      const float wm1 = (__uint_as_float(rm[8 * cc + 2 * e + 1]) + coef * vm[2 * e + 1]) * m1;
      s0 += wp0 + wm0;
      s1 += wp1 + wm1;
      du[cc * 4 + e] = pack_bf16x2(wp0 - wm0, wp1 - wm1);
    }
  }
  return s0 + s1;
}
// A warp's 32 half rows (64 bytes each, one per lane) leave through a warp-private 2 KB staging block: written row per
// lane, read back so that FOUR lanes cover one half row and one store instruction touches 8 lines instead of 32 -- the
// row-per-lane 16-byte stores of dq / dk / dv kept the load / store unit busy for ~3000 clocks of every 20 000-clock chunk
// (measured by switching them off: 756 -> 576 us per layer).  Chunk c of row l sits at l * 64 + ((c ^ ((l >> 1) & 3)) << 4):
// conflict-free both ways.  `base` = element (first row of the warp, column 32 g of the head); rows >= nvalid are skipped.
__device__ __forceinline__ void warp_rows_out(uint32_t stg, const uint4 (&t)[4], bf16* base, int64_t ld, int lane, int nvalid) {
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) sts128(stg + lane * 64 + (((uint32_t)cc ^ sw) << 4), t[cc]);
  __syncwarp();
  const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = 8 * j + sub;
    const uint4 v = lds128(stg + row * 64 + (((uint32_t)ch ^ (((uint32_t)row >> 1) & 3u)) << 4));
    if (row < nvalid) *(reinterpret_cast<uint4*>(base + (int64_t)row * ld) + ch) = v;
  }
  __syncwarp();
}
// The reverse for the q / k half rows the dx epilogues need: requested with the coalesced mapping (lane = (row & 7, chunk)
// of each group of 8 rows) a phase ahead, turned into "lane = row" through the staging block when they are used.
__device__ __forceinline__ void warp_rows_in_issue(uint4 (&t)[4], const bf16* base, int64_t ld, int lane, int nvalid) {
  const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = 8 * j + sub;
    t[j] = (row < nvalid) ? __ldg(reinterpret_cast<const uint4*>(base + (int64_t)row * ld) + ch) : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void warp_rows_in_finish(uint4 (&t)[4], uint32_t stg, int lane) {
  const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = 8 * j + sub;
    sts128(stg + row * 64 + (((uint32_t)ch ^ (((uint32_t)row >> 1) & 3u)) << 4), t[j]);
  }
  __syncwarp();
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) t[cc] = lds128(stg + lane * 64 + (((uint32_t)cc ^ sw) << 4));
  __syncwarp();
}
// columns [32 g, 32 g + 32) of a dx row = acc * ln2 + cx * x  -> global through the warp's staging block
__device__ __forceinline__ void dx_half_out(uint32_t ta, const uint4 (&xv)[4], uint32_t stg, bf16* base, int64_t ld, int lane, int nvalid,
                                            int g, float cx) {
  uint32_t r[32];
  tmem_ld32_issue(ta + g * 32, r);
  tmem_ld_wait();
  uint4 t[4];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const uint4 xx = xv[cc];
    float x0, x1;
    unpack_bf16x2(xx.x, x0, x1); t[cc].x = pack_bf16x2(__uint_as_float(r[8 * cc + 0]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 1]) * KINV + cx * x1);
    unpack_bf16x2(xx.y, x0, x1); t[cc].y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 3]) * KINV + cx * x1);
    unpack_bf16x2(xx.z, x0, x1); t[cc].z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 5]) * KINV + cx * x1);
    unpack_bf16x2(xx.w, x0, x1); t[cc].w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 7]) * KINV + cx * x1);
  }
  warp_rows_out(stg, t, base, ld, lane, nvalid);
}

__global__ void __launch_bounds__(NT, 1)
favor_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment (128B swizzle atoms) by pointer arithmetic ON the __shared__ array: an integer round trip would
  // turn every access through z / gd / sp / part into a generic load (ncu: LD.E with long-scoreboard stalls)
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(smem);
  const uint32_t sXQ = sb + OFF_XQ, sXK = sb + OFF_XK, sXV0 = sb + OFF_XV, sXO = sb + OFF_XO, sXD = sb + OFF_XD, sPQ = sb + OFF_PQ,
                 sPK = sb + OFF_PK, sG = sb + OFF_G, sSB = sb + OFF_SB, sRB = sb + OFF_RB, sOM = sb + OFF_OM;
  float* vec = reinterpret_cast<float*>(smem + OFF_VEC);
  float* z = vec + V_Z / 4;
  float* rzb = vec + V_RZ / 4;        // [2][128]
  float* gd = vec + V_GD / 4;
  float* sp = vec + V_SP / 4;         // [2 (q / k)][2 (g)][128]: partial sums of d phi * phi
  float* part = vec + V_PART / 4;     // [8][128]
  auto bar = [&](int i) { return sb + OFF_BAR + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ctrl = warp == 8;
  const int g = (warp >> 2) & 1, wq = warp & 3;

  if (tid == 256) {
    prefetch_map(&tmQ); prefetch_map(&tmK); prefetch_map(&tmV); prefetch_map(&tmO); prefetch_map(&tmD);
    for (int i = 0; i < N_BARS; ++i) mbar_init(bar(i), (i == F_X || i == F_OD || i == F_PHI) ? 256 : 1);
    mbar_init_fence();
  }
  if (ctrl) tmem_alloc(smem_u32(tmem_slot), T_COLS);
  stage_omega(p.omega, sOM, tid, NT, false);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tl = tmem + ((uint32_t)wq << 21);
  const int r = tid & 127;                      // TMEM lane: token row of every tile / feature row of the states

  constexpr uint32_t ID_U = make_idesc(64, false, true), ID_KK128 = make_idesc(128, false, false), ID_ST = make_idesc(64, true, true),
                     ID_MN64 = make_idesc(64, false, true), ID_KK64 = make_idesc(64, false, false);
  // Operand descriptors as (lo, hi) halves built from warp-uniform values OUTSIDE the elected-lane regions (they stay in
  // uniform registers; inside, a descriptor is one add of a constant): tc_ptx.cuh, elect_one.
  auto dK = [](uint32_t tile, int ks) { return desc_at(make_desc_lh(tile, 0, 1024), (uint32_t)(ks >> 2) * TILE + (uint32_t)(ks & 3) * 32); };   // K-major over 128 features
  auto d64 = [](uint32_t tile, int ks) { return desc_at(make_desc_lh(tile, 0, 1024), (uint32_t)ks * 32); };                                  // K-major, K <= 64
  auto dMN = [](uint32_t tile, int ks) { return desc_at(make_desc_lh(tile, TILE, 1024), (uint32_t)ks * 2048); };                             // MN-major, K = rows
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);      // the control warp's copy of the TMEM base, warp-uniform

  uint32_t cph = 0;            // parity of this chunk's once-per-chunk barriers
  uint32_t nchunks_done = 0;   // chunks processed by this CTA so far (V double buffer / rz double buffer index)

  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int bh = item / p.nseg, seg = item % p.nseg;
    const int b = bh / p.H, h = bh % p.H;
    const int nchunk = (p.T + C - 1) / C;
    const int c_begin = seg * p.seg_chunks;
    const int c_end = (c_begin + p.seg_chunks < nchunk) ? c_begin + p.seg_chunks : nchunk;
    if (c_begin >= c_end) continue;
    const int64_t tokbase = (int64_t)b * p.T;

    if (ctrl) {
      if (elect_one()) {
        const int t0 = (c_end - 1) * C;
        mbar_expect_tx(bar(B_XQK), 2 * TILE);
        tma_load_3d(&tmQ, bar(B_XQK), sXQ, h * FE, t0, b);
        tma_load_3d(&tmK, bar(B_XQK), sXK, h * FE, t0, b);
        mbar_expect_tx(bar(B_OD), 2 * TILE);
        tma_load_3d(&tmO, bar(B_OD), sXO, h * FE, t0, b);
        tma_load_3d(&tmD, bar(B_OD), sXD, h * FE, t0, b);
        const uint32_t vb = nchunks_done & 1;
        mbar_expect_tx(bar(B_V0 + vb), TILE);
        tma_load_3d(&tmV, bar(B_V0 + vb), sXV0 + vb * TILE, h * FE, t0, b);
      }
      __syncwarp();
    } else {
      // ---- states of this segment: group 0 loads R, rz (reverse state of the later segments), group 1 loads -S, z
      //      (the forward's prefix state at the end of the segment): fp32 -> TMEM, R also bf16 -> smem ----
      const float* src;
      if (g == 0) src = p.nseg > 1 ? p.seg_rstates + ((int64_t)bh * (p.nseg + 1) + seg) * FM * FV : nullptr;
      else {
        int slot = (seg + 1) * p.ratio;
        if (slot > p.fwd_nseg) slot = p.fwd_nseg;
        src = p.seg_states + ((int64_t)bh * (p.fwd_nseg + 1) + slot) * FM * FV;
      }
      const float* rp = src ? src + r * FV : nullptr;
      const float sg = g == 0 ? 1.f : -1.f;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = rp ? *reinterpret_cast<const float4*>(rp + half * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * j] = __float_as_uint(sg * t.x); v[4 * j + 1] = __float_as_uint(sg * t.y);
          v[4 * j + 2] = __float_as_uint(sg * t.z); v[4 * j + 3] = __float_as_uint(sg * t.w);
        }
        tmem_st32(tl + (g == 0 ? T_R : T_NS) + half * 32, v);
        if (g == 0) pack_row32(sRB, r, half * 4, v, 1.f);
      }
      if (g == 0) rzb[(nchunks_done & 1) * 128 + r] = rp ? rp[FE] : 0.f;
      else z[r] = rp ? rp[FE] : 0.f;
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
    }

    for (int c = c_end - 1; c >= c_begin; --c, cph ^= 1, ++nchunks_done) {
      const int t0 = c * C;
      const int valid = (p.T - t0 < C) ? (p.T - t0) : C;
      const uint32_t vb = nchunks_done & 1;
      const uint32_t sXV = sXV0 + vb * TILE;
      float* rz_cur = rzb + vb * 128;              // rz of the later chunks
      float* rz_nxt = rzb + (vb ^ 1) * 128;
      if (ctrl) {
        // =============================== control warp ===============================
        // all 32 lanes walk the protocol (waits, fences); one elected lane issues MMAs / TMA
        mbar_wait(bar(B_XQK), cph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_W1, d64(sXQ, ks), dMN(sOM, ks), ID_U, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_W1 + 64, d64(sXK, ks), dMN(sOM, ks), ID_U, ks > 0);
          umma_commit(bar(M_U));
          if (c > c_begin) {                       // prefetch the previous chunk (reverse order)
            mbar_expect_tx(bar(B_V0 + (vb ^ 1)), TILE);
            tma_load_3d(&tmV, bar(B_V0 + (vb ^ 1)), sXV0 + (vb ^ 1) * TILE, h * FE, t0 - C, b);
          }
        }
        __syncwarp();
        mbar_wait(bar(F_X), cph);
        if (c > c_begin && elect_one()) {
          mbar_expect_tx(bar(B_XQK), 2 * TILE);
          tma_load_3d(&tmQ, bar(B_XQK), sXQ, h * FE, t0 - C, b);
          tma_load_3d(&tmK, bar(B_XQK), sXK, h * FE, t0 - C, b);
        }
        __syncwarp();
        // phi(q), phi(k) are in shared memory (and every worker is past the previous chunk's last read of W0) while the
        // workers still form G / gd / the z roll-back: the first score tile is issued NOW, so that it is complete when
        // they come out of barrier [A] (they used to wait ~1400 clocks for it there: 778 -> 753 us per layer).
        // (Issuing the dq / dk products early the same way -- into the hole of W0, behind two more arrival barriers --
        // was built and measured: no change, those waits are not on the critical path.)
        mbar_wait(bar(F_PHI), cph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tm + T_W0, dK(sPK, ks), dK(sPQ, ks), ID_KK128, ks > 0);    // phi(k) phi(q)^T
          umma_commit(bar(M_C1));
        }
        __syncwarp();
        mbar_wait(bar(F_OD), cph);
        if (c > c_begin && elect_one()) {
          mbar_expect_tx(bar(B_OD), 2 * TILE);
          tma_load_3d(&tmO, bar(B_OD), sXO, h * FE, t0 - C, b);
          tma_load_3d(&tmD, bar(B_OD), sXD, h * FE, t0 - C, b);
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [A] G in smem; gd; z rolled back
        mbar_wait(bar(B_V0 + vb), (nchunks_done >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_W1, d64(sG, ks), d64(sXV, ks), ID_KK128, ks > 0);   // G V^T
          umma_commit(bar(M_C2));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_W2, d64(sXV, ks), d64(sG, ks), ID_KK128, ks > 0);   // V G^T
          umma_commit(bar(M_C3));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tm + T_NS, dMN(sPK, ks), dMN(sXV, ks), ID_ST, 1u);        // -S += phi(k)^T V
          umma_commit(bar(M_NS));
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [B] P^T, dP, dP^T packed in TMEM; S_prev bf16 in smem
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {         // d phi(q): features 64 hf .. -> hole of W1 / W0
            const uint32_t d = tmem + (hf == 0 ? T_W1 : T_W0) + 64;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(d, tm + T_W1 + ks * 8, dMN(sPK + hf * TILE, ks), ID_MN64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ss(d, d64(sG, ks), d64(sSB + hf * 8192, ks), ID_KK64, 1u);
          }
          umma_commit(bar(M_E1));
          {                                        // dv = P^T G + phi(k) R -> hole of W2
            const uint32_t d = tm + T_W2 + 64;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(d, tm + T_W0 + ks * 8, dMN(sG, ks), ID_MN64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ss(d, dK(sPK, ks), dMN(sRB, ks), ID_MN64, 1u);
          }
          umma_commit(bar(M_E2));
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [C] d phi(q) and dv consumed; dU_q packed in TMEM (W1 low)
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tm + T_W2 + 64, tm + T_W1 + ks * 8, d64(sOM, ks), ID_KK64, ks > 0);   // dU_q Om'^T
          umma_commit(bar(M_DXQ));
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {         // d phi(k)
            const uint32_t d = tmem + (hf == 0 ? T_W1 : T_W0) + 64;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(d, tm + T_W2 + ks * 8, dMN(sPQ + hf * TILE, ks), ID_MN64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ss(d, d64(sXV, ks), d64(sRB + hf * 8192, ks), ID_KK64, 1u);
          }
          umma_commit(bar(M_G));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tm + T_R, dMN(sPQ, ks), dMN(sG, ks), ID_ST, 1u);          // R += phi(q)^T G
          umma_commit(bar(M_R));
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [D] dx_q, d phi(k) consumed; dU_k packed in TMEM (W2 low); R bf16 in smem
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tm + T_W0 + 64, tm + T_W2 + ks * 8, d64(sOM, ks), ID_KK64, ks > 0);   // dU_k Om'^T
          umma_commit(bar(M_DXK));
        }
        __syncwarp();
      } else {
        // =============================== workers: thread (g, r) ===============================
        const bool rowok = r < valid;

        // output staging: S_prev's bf16 tile is dead between batch E1 of this chunk and phase 2 of the next one
        const uint32_t stg = sSB + (uint32_t)warp * 2048u;
        const int wrow0 = 32 * wq;                                  // first tile row of this warp
        const int nvalid_w = (p.debug & 1) ? 0 : (valid - wrow0 < 0 ? 0 : (valid - wrow0 > 32 ? 32 : valid - wrow0));
        const int64_t obase = (tokbase + t0 + wrow0) * p.ld_d + (int64_t)h * FE + 32 * g;
        const int64_t ibase = (tokbase + t0 + wrow0) * p.ld + (int64_t)h * FE + 32 * g;
        const int nvalid_in = (p.debug & 2) ? 0 : (valid - wrow0 < 0 ? 0 : (valid - wrow0 > 32 ? 32 : valid - wrow0));
        const int64_t tok = tokbase + t0 + r;
        // 1 / den of this row, requested before anything waits
        const float den_r = rowok ? __ldg(p.den + tok * p.H + h) : 1.f;
        // ---- phase 1: phi of this group's tensor (g = 0: q, g = 1: k) -> smem; G, gd; z rolled back ----
        mbar_wait(bar(M_U), cph);
        tc_fence_after();
        {
          const float ss = row_sumsq(g ? sXK : sXQ, r);
          mbar_arrive_after_reads(bar(F_X), ss);
          const float o = rowok ? (0.5f * F_S2 * ss + F_HALF_LOG_M) * K2 : __int_as_float(0x7f800000);   // +inf -> phi = 0
          phi_row_to_smem(tl + T_W1 + 64 * g, g ? sPK : sPQ, r, o);
          fence_proxy_async();                     // this row of phi is visible to the tensor core
          tc_fence_before();
          mbar_arrive(bar(F_PHI));
        }
        {
          mbar_wait(bar(B_OD), cph);
          const float inv = rowok ? 1.f / den_r : 0.f;
          if (g == 0) {                            // the normaliser's gradient: gd = -(dout . out) / den
            float dot = 0.f;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const uint4 od = lds128(sXO + sw128(r, cc)), dd = lds128(sXD + sw128(r, cc));
              float o0, o1, d0, d1;
              unpack_bf16x2(od.x, o0, o1); unpack_bf16x2(dd.x, d0, d1); dot += o0 * d0 + o1 * d1;
              unpack_bf16x2(od.y, o0, o1); unpack_bf16x2(dd.y, d0, d1); dot += o0 * d0 + o1 * d1;
              unpack_bf16x2(od.z, o0, o1); unpack_bf16x2(dd.z, d0, d1); dot += o0 * d0 + o1 * d1;
              unpack_bf16x2(od.w, o0, o1); unpack_bf16x2(dd.w, d0, d1); dot += o0 * d0 + o1 * d1;
            }
            gd[r] = -dot * inv;
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {         // G = dout / den, columns [32 g, 32 g + 32)
            const uint4 dd = lds128(sXD + sw128(r, g * 4 + cc));
            float d0, d1;
            uint4 t;
            unpack_bf16x2(dd.x, d0, d1); t.x = pack_bf16x2(d0 * inv, d1 * inv);
            unpack_bf16x2(dd.y, d0, d1); t.y = pack_bf16x2(d0 * inv, d1 * inv);
            unpack_bf16x2(dd.z, d0, d1); t.z = pack_bf16x2(d0 * inv, d1 * inv);
            unpack_bf16x2(dd.w, d0, d1); t.w = pack_bf16x2(d0 * inv, d1 * inv);
            sts128(sG + sw128(r, g * 4 + cc), t);
          }
          mbar_arrive(bar(F_OD));
        }
        if (g == 1) {
          named_bar_sync<2>(128);                  // phi(k) tile complete
          colsum_partial<128>(sPK, r, nullptr, part);
          named_bar_sync<2>(128);
          z[r] -= part_total(part, r);
        }
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [A]
        const float gdr = gd[r];
        // ---- phase 2: the three masked tiles, packed in place; S_prev -> bf16 smem ----
        mbar_wait(bar(M_C1), cph);
        tc_fence_after();
        mask_tile<0, 3>(tl + T_W0, g, wq, r, 0.f, gd);
        mbar_wait(bar(M_C2), cph);
        tc_fence_after();
        mask_tile<1, 4>(tl + T_W1, g, wq, r, gdr, gd);
        mbar_wait(bar(M_C3), cph);
        tc_fence_after();
        mask_tile<2, 5>(tl + T_W2, g, wq, r, 0.f, gd);
        mbar_wait(bar(M_NS), cph);
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld32_issue(tl + T_NS + g * 32, v);
          tmem_ld_wait();
          pack_row32(sSB, r, g * 4, v, -1.f);      // feature row r of S_prev = -(-S)
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [B]
        // ---- phase 3: rz of the next (earlier) chunk; d phi(q) -> dU_q; dv out ----
        uint4 xq4[4];                              // requested a phase ahead of its use
        warp_rows_in_issue(xq4, p.q + ibase, p.ld, lane, nvalid_in);
        colsum_partial<256>(sPQ, tid, gd, part);   // sum_i gd_i phi(q_i)
        named_bar_sync<6>(256);
        if (g == 0) rz_nxt[r] = rz_cur[r] + part_total(part, r);
        {
          uint32_t du[16];
          mbar_wait(bar(M_E1), cph);
          tc_fence_after();
          sp[g * 128 + r] = dphi_half(tl + T_W1 + 64, tl + T_W0 + 64, sPQ, r, g, z, gdr, du);
          tmem_st16(tl + T_W1 + g * 16, du);       // dP (W1 low) is dead: batch E1 has completed
        }
        mbar_wait(bar(M_E2), cph);
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld32_issue(tl + T_W2 + 64 + g * 32, v);
          tmem_ld_wait();
          uint4 t[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            t[cc].x = pack_bf16x2(__uint_as_float(v[8 * cc]), __uint_as_float(v[8 * cc + 1]));
            t[cc].y = pack_bf16x2(__uint_as_float(v[8 * cc + 2]), __uint_as_float(v[8 * cc + 3]));
            t[cc].z = pack_bf16x2(__uint_as_float(v[8 * cc + 4]), __uint_as_float(v[8 * cc + 5]));
            t[cc].w = pack_bf16x2(__uint_as_float(v[8 * cc + 6]), __uint_as_float(v[8 * cc + 7]));
          }
          warp_rows_out(stg, t, p.dv + obase, p.ld_d, lane, nvalid_w);
        }
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [C]
        // ---- phase 4: dq out; d phi(k) -> dU_k; R -> bf16 smem ----
        uint4 xk4[4];
        warp_rows_in_issue(xk4, p.k + ibase, p.ld, lane, nvalid_in);
        mbar_wait(bar(M_DXQ), cph);
        tc_fence_after();
        warp_rows_in_finish(xq4, stg, lane);
        dx_half_out(tl + T_W2 + 64, xq4, stg, p.dq + obase, p.ld_d, lane, nvalid_w, g, -(sp[r] + sp[128 + r]) * F_S2);
        {
          uint32_t du[16];
          mbar_wait(bar(M_G), cph);
          tc_fence_after();
          sp[256 + g * 128 + r] = dphi_half(tl + T_W1 + 64, tl + T_W0 + 64, sPK, r, g, rz_cur, 1.f, du);
          tmem_st16(tl + T_W2 + g * 16, du);       // dP^T (W2 low) is dead: batch G has completed
        }
        mbar_wait(bar(M_R), cph);                  // every MMA of this chunk that reads R bf16 / phi(q) / G has completed
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld32_issue(tl + T_R + g * 32, v);
          tmem_ld_wait();
          pack_row32(sRB, r, g * 4, v, 1.f);
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [D]
        // ---- phase 5: dk out ----
        mbar_wait(bar(M_DXK), cph);
        tc_fence_after();
        warp_rows_in_finish(xk4, stg, lane);
        dx_half_out(tl + T_W0 + 64, xk4, stg, p.dk + obase, p.ld_d, lane, nvalid_w, g, -(sp[256 + r] + sp[384 + r]) * F_S2);
        tc_fence_before();
      }
    }
    // all roles meet before the next item rewrites the states
    tc_fence_before();
    named_bar_sync<1>(NT);
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (ctrl) {
    tc_fence_after();
    tmem_dealloc(tmem, T_COLS);
  }
}

}  // namespace favor3b

// nseg segments of sc 128-token chunks per (b, h); workspace arguments as emo_favor_bwd
int emo_favor_bwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, const void* out,
                            const void* dout, int64_t ld_out, const float* den, const float* seg_states,
                            const float* seg_rstates, int nseg, int sc, int fwd_nseg, int ratio, void* dq, void* dk, void* dv,
                            int64_t ld_d, int B, int T_, int H, cudaStream_t s) {
  using namespace favor3b;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  CUtensorMap mq, mk, mv, mo, md;
  int rc;
  if ((rc = make_map_bt(&mq, q, (int64_t)H * FE, T_, B, ld, C))) return rc;
  if ((rc = make_map_bt(&mk, k, (int64_t)H * FE, T_, B, ld, C))) return rc;
  if ((rc = make_map_bt(&mv, v, (int64_t)H * FE, T_, B, ld, C))) return rc;
  if ((rc = make_map_bt(&mo, out, (int64_t)H * FE, T_, B, ld_out, C))) return rc;
  if ((rc = make_map_bt(&md, dout, (int64_t)H * FE, T_, B, ld_out, C))) return rc;
  Params p;
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.ld = ld; p.omega = omega; p.den = den; p.seg_states = seg_states;
  p.seg_rstates = seg_rstates; p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.ld_d = ld_d;
  { const char* e = getenv("EMO_FAVOR_BWD_DBG"); p.debug = e ? atoi(e) : 0; }
  p.nseg = nseg; p.seg_chunks = sc; p.fwd_nseg = fwd_nseg; p.ratio = ratio; p.T = T_; p.H = H; p.items = B * H * nseg;
  const int sms = emo_num_sms();
  const int grid = p.items < sms ? p.items : sms;
  favor_bwd_tc_kernel<<<grid, NT, SMEM_BYTES, s>>>(mq, mk, mv, mo, md, p);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
