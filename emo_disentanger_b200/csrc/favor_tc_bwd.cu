// FAVOR+ backward on the 5th-gen tensor cores (bf16, sm_100a): every product of the reverse pass is a tcgen05.mma with
// its accumulator in TMEM; q / k / v / out / dout chunks are staged by TMA (128B swizzle).  Same math, interface and
// workspace layout as favor_bwd2_kernel (mma.sync), which stays as the A/B switch EMO_FAVOR_TC=0.
//
// Backward of fast_transformers CausalLinearAttention + Favor (causal_dot_product's backward and the feature map's),
// stage2_accompaniment/model/fast_transformer_decoder.py:28-38.
//
// One CTA per SM walks one (batch, head, segment) in REVERSE in chunks of 128 tokens.  8 worker warps in two groups --
// the query side (warps 0-3, thread = query row i = TMEM lane) and the key side (warps 4-7, thread = key row j) --
// and one control warp (TMA + MMA issue).  With G = dout / den, gd = -(dout . out) / den (the normaliser's gradient),
// P = phi(q) phi(k)^T and dP_ij = G_i . v_j + gd_i (both causal), per chunk:
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
#include "tc_ptx.cuh"

namespace favor3b {
using namespace tcp;

constexpr int C = 128, FE = 64, FM = 128, FV = 80, NT = 288;
constexpr float F_S2 = 0.125f, F_HALF_LOG_M = 2.4260151319598084f, K2 = 1.4426950408889634f, KINV = 0.6931471805599453f;
constexpr uint32_t TILE = 16384;
constexpr uint32_t OFF_XQ = 0, OFF_XK = TILE, OFF_XV = 2 * TILE /* x2 */, OFF_XO = 4 * TILE, OFF_XD = 5 * TILE, OFF_PQ = 6 * TILE /* x2 */,
                   OFF_PK = 8 * TILE /* x2 */, OFF_G = 10 * TILE, OFF_SB = 11 * TILE, OFF_RB = 12 * TILE, OFF_OM = 13 * TILE /* 8 KB */,
                   OFF_VEC = 13 * TILE + 8192;
// fp32 vectors: z[128] rz[2][128] gd[128] part[8][128]
constexpr uint32_t V_Z = 0, V_RZ = 512, V_GD = 1536, V_PART = 2048, VEC_BYTES = 2048 + 4096;
constexpr uint32_t OFF_BAR = OFF_VEC + VEC_BYTES, SMEM_USED = OFF_BAR + 256;
constexpr int SMEM_BYTES = SMEM_USED + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");
constexpr uint32_t T_R = 0, T_NS = 64, T_W1 = 128, T_W0 = 256, T_W2 = 384, T_COLS = 512;

// barriers (8 bytes each, from OFF_BAR)
enum { B_XQK = 0, B_OD, B_V0, B_V1, F_X, F_OD, M_U, M_NS, M_C2, M_C13, M_E1, M_E2, M_G, M_DXQ, M_R, M_DXK, N_BARS };

struct Params {
  const bf16* q; const bf16* k; int64_t ld;
  const float* omega; const float* den;
  const float* seg_states; const float* seg_rstates;
  bf16* dq; bf16* dk; bf16* dv; int64_t ld_d;
  int nseg, seg_chunks, fwd_nseg, ratio, T, H, items, omega_f16;
};

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
#pragma unroll
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
// column sums over the 128 token rows of a [token][128 features] tile, optionally weighted per row: 128 threads,
// thread t -> 16-byte chunk column (t & 15) of the two blocks, 16 rows (t >> 4); partial sums to part[8][128]
__device__ __forceinline__ void colsum_partial(uint32_t tile, int t, const float* w, float* part) {
  const int cc = t & 15, blk = cc >> 3, c = cc & 7, rg = t >> 4;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
  for (int r = rg * 16; r < rg * 16 + 16; ++r) {
    const uint4 a = lds128(tile + blk * TILE + sw128(r, c));
    const float wr = w ? w[r] : 1.f;
    float f0, f1;
    unpack_bf16x2(a.x, f0, f1); acc[0] += wr * f0; acc[1] += wr * f1;
    unpack_bf16x2(a.y, f0, f1); acc[2] += wr * f0; acc[3] += wr * f1;
    unpack_bf16x2(a.z, f0, f1); acc[4] += wr * f0; acc[5] += wr * f1;
    unpack_bf16x2(a.w, f0, f1); acc[6] += wr * f0; acc[7] += wr * f1;
  }
  float4* dst = reinterpret_cast<float4*>(part + rg * 128 + blk * 64 + c * 8);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}
__device__ __forceinline__ float part_total(const float* part, int f) {
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) s += part[g * 128 + f];
  return s;
}

// d phi (two 64-column accumulators: features 0..63 at tp, 64..127 at tm) of this thread's row -> du (packed bf16,
// 32 words) and the coefficient of x:   w = (d phi + coef * vec) * phi ;  du = w+ - w- ;  cx = -sum(w) * s^2
__device__ __forceinline__ void dphi_row(uint32_t tp, uint32_t tm, uint32_t phi_tile, int row, const float* vec, float coef,
                                         uint32_t (&du)[32], float& cx) {
  float s0 = 0.f, s1 = 0.f;
  const float4* v4 = reinterpret_cast<const float4*>(vec);
#pragma unroll
  for (int half = 0; half < 2; ++half) {          // features 32 half .. 32 half + 31 of both signs
    uint32_t rp[32], rm[32];
    tmem_ld32_issue(tp + half * 32, rp);
    tmem_ld32_issue(tm + half * 32, rm);
    tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const uint4 php = lds128(phi_tile + sw128(row, half * 4 + cc)), phm = lds128(phi_tile + TILE + sw128(row, half * 4 + cc));
      const float4 vpa = v4[half * 8 + 2 * cc], vpb = v4[half * 8 + 2 * cc + 1];
      const float4 vma = v4[16 + half * 8 + 2 * cc], vmb = v4[16 + half * 8 + 2 * cc + 1];
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
        du[half * 16 + cc * 4 + e] = pack_bf16x2(wp0 - wm0, wp1 - wm1);
      }
    }
  }
  cx = -(s0 + s1) * F_S2;
}
// dx row = acc(64 columns at `ta`) * ln2 + cx * x  -> global (16-byte stores); x row re-read from global (L2)
__device__ __forceinline__ void dx_row_out(uint32_t ta, const bf16* xrow, bf16* drow, float cx, bool ok) {
  uint4 xv[8];
  if (ok) {
#pragma unroll
    for (int c = 0; c < 8; ++c) xv[c] = __ldg(reinterpret_cast<const uint4*>(xrow) + c);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld32_issue(ta + half * 32, r);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const uint4 xx = xv[half * 4 + cc];
        float x0, x1;
        uint4 t;
        unpack_bf16x2(xx.x, x0, x1); t.x = pack_bf16x2(__uint_as_float(r[8 * cc + 0]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 1]) * KINV + cx * x1);
        unpack_bf16x2(xx.y, x0, x1); t.y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 3]) * KINV + cx * x1);
        unpack_bf16x2(xx.z, x0, x1); t.z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 5]) * KINV + cx * x1);
        unpack_bf16x2(xx.w, x0, x1); t.w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]) * KINV + cx * x0, __uint_as_float(r[8 * cc + 7]) * KINV + cx * x1);
        *(reinterpret_cast<uint4*>(drow) + half * 4 + cc) = t;
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 1)
favor_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  const uint32_t sXQ = sb + OFF_XQ, sXK = sb + OFF_XK, sXV0 = sb + OFF_XV, sXO = sb + OFF_XO, sXD = sb + OFF_XD, sPQ = sb + OFF_PQ,
                 sPK = sb + OFF_PK, sG = sb + OFF_G, sSB = sb + OFF_SB, sRB = sb + OFF_RB, sOM = sb + OFF_OM;
  float* vec = reinterpret_cast<float*>(smem + OFF_VEC);
  float* z = vec + V_Z / 4;
  float* rzb = vec + V_RZ / 4;        // [2][128]
  float* gd = vec + V_GD / 4;
  float* part = vec + V_PART / 4;     // [8][128]
  auto bar = [&](int i) { return sb + OFF_BAR + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ctrl = warp == 8, qside = warp < 4;

  if (tid == 256) {
    prefetch_map(&tmQ); prefetch_map(&tmK); prefetch_map(&tmV); prefetch_map(&tmO); prefetch_map(&tmD);
    for (int i = 0; i < N_BARS; ++i) mbar_init(bar(i), i == F_X ? 256 : (i == F_OD ? 128 : 1));
    mbar_init_fence();
  }
  if (ctrl) tmem_alloc(smem_u32(tmem_slot), T_COLS);
  stage_omega(p.omega, sOM, tid, NT, p.omega_f16 != 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tl = tmem + ((uint32_t)(warp & 3) << 21);
  const int row = tid & 127;                    // token row of the chunk (query side: i, key side: j) / feature row f

  constexpr uint32_t ID_KK128 = make_idesc(128, false, false), ID_ST = make_idesc(64, true, true), ID_MN64 = make_idesc(64, false, true),
                     ID_KK64 = make_idesc(64, false, false);
  // products against Omega' (fp16 tile when p.omega_f16): U = X Om' and dx = dU Om'^T
  const uint32_t ID_U = p.omega_f16 ? make_idesc(64, false, true, 128, true, false) : make_idesc(64, false, true);
  const uint32_t ID_DX = p.omega_f16 ? make_idesc(64, false, false, 128, true, false) : ID_KK64;
  auto dK = [](uint32_t tile, int ks) { return make_desc(tile + (uint32_t)(ks >> 2) * TILE + (uint32_t)(ks & 3) * 32, 0, 1024); };   // K-major over 128 features
  auto d64 = [](uint32_t tile, int ks) { return make_desc(tile + (uint32_t)ks * 32, 0, 1024); };                                  // K-major, K <= 64
  auto dMN = [](uint32_t tile, int ks) { return make_desc(tile + (uint32_t)ks * 2048, TILE, 1024); };                             // MN-major, K = rows

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
      if (lane == 0) {
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
    } else {
      // ---- states of this segment: R, rz (query side) and -S, z (key side): fp32 -> TMEM, R also bf16 -> smem ----
      const float* src;
      if (qside) src = p.nseg > 1 ? p.seg_rstates + ((int64_t)bh * (p.nseg + 1) + seg) * FM * FV : nullptr;
      else {
        int slot = (seg + 1) * p.ratio;
        if (slot > p.fwd_nseg) slot = p.fwd_nseg;
        src = p.seg_states + ((int64_t)bh * (p.fwd_nseg + 1) + slot) * FM * FV;
      }
      const float* rp = src ? src + row * FV : nullptr;
      const float sg = qside ? 1.f : -1.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = rp ? *reinterpret_cast<const float4*>(rp + half * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          r[4 * j] = __float_as_uint(sg * t.x); r[4 * j + 1] = __float_as_uint(sg * t.y);
          r[4 * j + 2] = __float_as_uint(sg * t.z); r[4 * j + 3] = __float_as_uint(sg * t.w);
        }
        tmem_st32(tl + (qside ? T_R : T_NS) + half * 32, r);
        if (qside) pack_row32(sRB, row, half * 4, r, 1.f);
      }
      if (qside) rzb[(nchunks_done & 1) * 128 + row] = rp ? rp[FE] : 0.f;
      else z[row] = rp ? rp[FE] : 0.f;
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
    }

    for (int c = c_end - 1; c >= c_begin; --c, cph ^= 1, ++nchunks_done) {
      const int t0 = c * C;
      const int valid = (p.T - t0 < C) ? (p.T - t0) : C;
      const uint32_t vb = nchunks_done & 1;
      const uint32_t sXV = sXV0 + vb * TILE;
      float* rz_cur = rzb + vb * 128;              // rz of the later chunks (read by the key side in phase 4)
      float* rz_nxt = rzb + (vb ^ 1) * 128;
      if (ctrl) {
        // =============================== control warp ===============================
        if (lane == 0) {
          mbar_wait(bar(B_XQK), cph);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tmem + T_W1, d64(sXQ, ks), dMN(sOM, ks), ID_U, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tmem + T_W1 + 64, d64(sXK, ks), dMN(sOM, ks), ID_U, ks > 0);
          umma_commit(bar(M_U));
          if (c > c_begin) {                       // prefetch the previous chunk (reverse order)
            const int tn = t0 - C;
            mbar_expect_tx(bar(B_V0 + (vb ^ 1)), TILE);
            tma_load_3d(&tmV, bar(B_V0 + (vb ^ 1)), sXV0 + (vb ^ 1) * TILE, h * FE, tn, b);
            mbar_wait(bar(F_X), cph);
            mbar_expect_tx(bar(B_XQK), 2 * TILE);
            tma_load_3d(&tmQ, bar(B_XQK), sXQ, h * FE, tn, b);
            tma_load_3d(&tmK, bar(B_XQK), sXK, h * FE, tn, b);
            mbar_wait(bar(F_OD), cph);
            mbar_expect_tx(bar(B_OD), 2 * TILE);
            tma_load_3d(&tmO, bar(B_OD), sXO, h * FE, tn, b);
            tma_load_3d(&tmD, bar(B_OD), sXD, h * FE, tn, b);
          } else {
            mbar_wait(bar(F_X), cph);
            mbar_wait(bar(F_OD), cph);
          }
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [A] phi(q), phi(k), G in smem; z rolled back
        if (lane == 0) {
          tc_fence_after();
          mbar_wait(bar(B_V0 + vb), (nchunks_done >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tmem + T_NS, dMN(sPK, ks), dMN(sXV, ks), ID_ST, 1u);        // -S += phi(k)^T V
          umma_commit(bar(M_NS));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tmem + T_W1, d64(sG, ks), d64(sXV, ks), ID_KK128, ks > 0);   // G V^T
          umma_commit(bar(M_C2));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tmem + T_W0, dK(sPK, ks), dK(sPQ, ks), ID_KK128, ks > 0);    // phi(k) phi(q)^T
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tmem + T_W2, d64(sXV, ks), d64(sG, ks), ID_KK128, ks > 0);   // V G^T
          umma_commit(bar(M_C13));
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [B] dP, P^T, dP^T packed in TMEM; S_prev bf16 in smem
        if (lane == 0) {
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {         // d phi(q): features 64 hf .. -> hole of W1 / W0
            const uint32_t d = tmem + (hf == 0 ? T_W1 : T_W0) + 64;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(d, tmem + T_W1 + ks * 8, dMN(sPK + hf * TILE, ks), ID_MN64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ss(d, d64(sG, ks), d64(sSB + hf * 8192, ks), ID_KK64, 1u);
          }
          umma_commit(bar(M_E1));
          {                                        // dv = P^T G + phi(k) R -> hole of W2
            const uint32_t d = tmem + T_W2 + 64;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(d, tmem + T_W0 + ks * 8, dMN(sG, ks), ID_MN64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ss(d, dK(sPK, ks), dMN(sRB, ks), ID_MN64, 1u);
          }
          umma_commit(bar(M_E2));
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [C] d phi(q) and dv consumed; dU_q packed in TMEM (W1 low)
        if (lane == 0) {
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {         // d phi(k)
            const uint32_t d = tmem + (hf == 0 ? T_W1 : T_W0) + 64;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(d, tmem + T_W2 + ks * 8, dMN(sPQ + hf * TILE, ks), ID_MN64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ss(d, d64(sXV, ks), d64(sRB + hf * 8192, ks), ID_KK64, 1u);
          }
          umma_commit(bar(M_G));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tmem + T_W2 + 64, tmem + T_W1 + ks * 8, d64(sOM, ks), ID_DX, ks > 0);   // dU_q Om'^T
          umma_commit(bar(M_DXQ));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tmem + T_R, dMN(sPQ, ks), dMN(sG, ks), ID_ST, 1u);          // R += phi(q)^T G
          umma_commit(bar(M_R));
        }
        __syncwarp();
        named_bar_sync<1>(NT);                     // [D] d phi(k), dx_q consumed; dU_k packed in TMEM (W2 low); R bf16 in smem
        if (lane == 0) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tmem + T_W0 + 64, tmem + T_W2 + ks * 8, d64(sOM, ks), ID_DX, ks > 0);   // dU_k Om'^T
          umma_commit(bar(M_DXK));
        }
        __syncwarp();
      } else if (qside) {
        // =============================== query side (row i) ===============================
        const int i = row;
        const bool rowok = i < valid;
        // ---- phase 1: phi(q) -> smem; G, gd ----
        mbar_wait(bar(M_U), cph);
        tc_fence_after();
        const float ssq = row_sumsq(sXQ, i);
        mbar_arrive(bar(F_X));
        const float oq = rowok ? (0.5f * F_S2 * ssq + F_HALF_LOG_M) * K2 : __int_as_float(0x7f800000);   // +inf -> phi = 0
        phi_row_to_smem(tl + T_W1, sPQ, i, oq);
        float gdi;
        {
          mbar_wait(bar(B_OD), cph);
          const float dn = rowok ? p.den[(tokbase + t0 + i) * p.H + h] : 1.f;
          const float inv = rowok ? 1.f / dn : 0.f;
          float dot = 0.f;
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const uint4 od = lds128(sXO + sw128(i, cc)), dd = lds128(sXD + sw128(i, cc));
            float o0, o1, d0, d1;
            uint4 t;
            unpack_bf16x2(od.x, o0, o1); unpack_bf16x2(dd.x, d0, d1); dot += o0 * d0 + o1 * d1; t.x = pack_bf16x2(d0 * inv, d1 * inv);
            unpack_bf16x2(od.y, o0, o1); unpack_bf16x2(dd.y, d0, d1); dot += o0 * d0 + o1 * d1; t.y = pack_bf16x2(d0 * inv, d1 * inv);
            unpack_bf16x2(od.z, o0, o1); unpack_bf16x2(dd.z, d0, d1); dot += o0 * d0 + o1 * d1; t.z = pack_bf16x2(d0 * inv, d1 * inv);
            unpack_bf16x2(od.w, o0, o1); unpack_bf16x2(dd.w, d0, d1); dot += o0 * d0 + o1 * d1; t.w = pack_bf16x2(d0 * inv, d1 * inv);
            sts128(sG + sw128(i, cc), t);
          }
          mbar_arrive(bar(F_OD));
          gdi = -dot * inv;
          gd[i] = gdi;
        }
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [A]
        // ---- phase 2: S_prev -> bf16 smem (feature row f = this thread); dP = tril(G V^T + gd_i) packed in place ----
        mbar_wait(bar(M_NS), cph);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld32_issue(tl + T_NS + half * 32, r);
          tmem_ld_wait();
          pack_row32(sSB, row, half * 4, r, -1.f);
        }
        mbar_wait(bar(M_C2), cph);
        tc_fence_after();
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          uint32_t pk[16];
          if (pc <= warp) {                        // key columns 32 pc .. : all in the future of rows < 32 pc
            uint32_t r[32];
            tmem_ld32_issue(tl + T_W1 + pc * 32, r);
            tmem_ld_wait();
            if (pc < warp) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]) + gdi, __uint_as_float(r[2 * j + 1]) + gdi);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int c0 = pc * 32 + 2 * j;
                pk[j] = pack_bf16x2(c0 <= i ? __uint_as_float(r[2 * j]) + gdi : 0.f, c0 + 1 <= i ? __uint_as_float(r[2 * j + 1]) + gdi : 0.f);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          tmem_st16(tl + T_W1 + pc * 16, pk);
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [B]
        // ---- phase 3: d phi(q) -> dU_q (packed, TMEM W1 low 32 columns), coefficient of x_q ----
        float cxq;
        {
          uint32_t du[32];
          mbar_wait(bar(M_E1), cph);
          tc_fence_after();
          dphi_row(tl + T_W1 + 64, tl + T_W0 + 64, sPQ, i, z, gdi, du, cxq);
          tmem_st32(tl + T_W1, du);                // dP (W1 low) is dead: batch E1 has completed
          tmem_st_wait();
        }
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [C]
        // ---- phase 4: dq; R -> bf16 smem ----
        mbar_wait(bar(M_DXQ), cph);
        tc_fence_after();
        dx_row_out(tl + T_W2 + 64, p.q + (tokbase + t0 + i) * p.ld + (int64_t)h * FE, p.dq + (tokbase + t0 + i) * p.ld_d + (int64_t)h * FE, cxq, rowok);
        mbar_wait(bar(M_R), cph);                  // every MMA of this chunk that reads R bf16 / phi(q) / G has completed
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld32_issue(tl + T_R + half * 32, r);
          tmem_ld_wait();
          pack_row32(sRB, row, half * 4, r, 1.f);
        }
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [D]
      } else {
        // =============================== key side (row j) ===============================
        const int j = row;
        const bool rowok = j < valid;
        // ---- phase 1: phi(k) -> smem; z rolled back to the start of the chunk ----
        mbar_wait(bar(M_U), cph);
        tc_fence_after();
        const float ssk = row_sumsq(sXK, j);
        mbar_arrive(bar(F_X));
        const float ok = rowok ? (0.5f * F_S2 * ssk + F_HALF_LOG_M) * K2 : __int_as_float(0x7f800000);
        phi_row_to_smem(tl + T_W1 + 64, sPK, j, ok);
        named_bar_sync<2>(128);                    // phi(k) tile complete (key side only)
        colsum_partial(sPK, row, nullptr, part);
        named_bar_sync<2>(128);
        z[row] -= part_total(part, row);
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [A]
        // ---- phase 2: P^T = triu(phi(k) phi(q)^T), dP^T = triu(V G^T + gd_i) packed in place ----
        mbar_wait(bar(M_C13), cph);
        tc_fence_after();
        const int kw = warp - 4;
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          uint32_t pk[16];
          if (pc >= kw) {                          // query columns 32 pc .. : all in the past of rows >= 32 (pc + 1)
            uint32_t r[32];
            tmem_ld32_issue(tl + T_W0 + pc * 32, r);
            tmem_ld_wait();
            if (pc > kw) {
#pragma unroll
              for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int c0 = pc * 32 + 2 * q;
                pk[q] = pack_bf16x2(c0 >= j ? __uint_as_float(r[2 * q]) : 0.f, c0 + 1 >= j ? __uint_as_float(r[2 * q + 1]) : 0.f);
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) pk[q] = 0u;
          }
          tmem_st16(tl + T_W0 + pc * 16, pk);
        }
        {
          const float4* g4 = reinterpret_cast<const float4*>(gd);
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t pk[16];
            if (pc >= kw) {
              uint32_t r[32];
              tmem_ld32_issue(tl + T_W2 + pc * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 gq = g4[pc * 8 + q];
                const int c0 = pc * 32 + 4 * q;
                float a0 = __uint_as_float(r[4 * q]) + gq.x, a1 = __uint_as_float(r[4 * q + 1]) + gq.y;
                float a2 = __uint_as_float(r[4 * q + 2]) + gq.z, a3 = __uint_as_float(r[4 * q + 3]) + gq.w;
                if (pc == kw) {
                  a0 = c0 >= j ? a0 : 0.f; a1 = c0 + 1 >= j ? a1 : 0.f; a2 = c0 + 2 >= j ? a2 : 0.f; a3 = c0 + 3 >= j ? a3 : 0.f;
                }
                pk[2 * q] = pack_bf16x2(a0, a1);
                pk[2 * q + 1] = pack_bf16x2(a2, a3);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) pk[q] = 0u;
            }
            tmem_st16(tl + T_W2 + pc * 16, pk);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [B]
        // ---- phase 3: dv out; rz of the next (earlier) chunk ----
        colsum_partial(sPQ, row, gd, part);        // sum_i gd_i phi(q_i): phi(q), gd were complete at [A]
        named_bar_sync<2>(128);
        rz_nxt[row] = rz_cur[row] + part_total(part, row);
        mbar_wait(bar(M_E2), cph);
        tc_fence_after();
        {
          bf16* drow = p.dv + (tokbase + t0 + j) * p.ld_d + (int64_t)h * FE;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld32_issue(tl + T_W2 + 64 + half * 32, r);
            tmem_ld_wait();
            if (rowok) {
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                uint4 t;
                t.x = pack_bf16x2(__uint_as_float(r[8 * cc]), __uint_as_float(r[8 * cc + 1]));
                t.y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]), __uint_as_float(r[8 * cc + 3]));
                t.z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]), __uint_as_float(r[8 * cc + 5]));
                t.w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]), __uint_as_float(r[8 * cc + 7]));
                *(reinterpret_cast<uint4*>(drow) + half * 4 + cc) = t;
              }
            }
          }
        }
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [C]
        // ---- phase 4: d phi(k) -> dU_k (packed, TMEM W2 low 32 columns), coefficient of x_k ----
        float cxk;
        {
          uint32_t du[32];
          mbar_wait(bar(M_G), cph);
          tc_fence_after();
          dphi_row(tl + T_W1 + 64, tl + T_W0 + 64, sPK, j, rz_cur, 1.f, du, cxk);
          tmem_st32(tl + T_W2, du);                // dP^T (W2 low) is dead: batch G has completed
          tmem_st_wait();
        }
        tc_fence_before();
        named_bar_sync<1>(NT);                     // [D]
        // ---- phase 5: dk ----
        mbar_wait(bar(M_DXK), cph);
        tc_fence_after();
        dx_row_out(tl + T_W0 + 64, p.k + (tokbase + t0 + j) * p.ld + (int64_t)h * FE, p.dk + (tokbase + t0 + j) * p.ld_d + (int64_t)h * FE, cxk, rowok);
        tc_fence_before();
      }
    }
    // all three roles meet before the next item rewrites the states (the key side may still be storing dk)
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
  p.nseg = nseg; p.seg_chunks = sc; p.fwd_nseg = fwd_nseg; p.ratio = ratio; p.T = T_; p.H = H; p.items = B * H * nseg; p.omega_f16 = favor_omega_f16();
  const int sms = emo_num_sms();
  const int grid = p.items < sms ? p.items : sms;
  favor_bwd_tc_kernel<<<grid, NT, SMEM_BYTES, s>>>(mq, mk, mv, mo, md, p);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
