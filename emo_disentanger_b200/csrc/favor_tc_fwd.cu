// FAVOR+ forward on the 5th-gen tensor cores (bf16, sm_100a): tcgen05.mma with fp32 accumulators in TMEM, q / k / v
// chunks staged by TMA (cp.async.bulk.tensor, 128B swizzle).  Same math, interface and workspace layout as
// favor_fwd2_kernel (mma.sync), which stays as the A/B switch EMO_FAVOR_TC=0.
//
// Replaces fast_transformers Favor.forward + CausalLinearAttention.forward + causal_dot_product
// (stage2_accompaniment/model/fast_transformer_decoder.py:28-38).
//
// One CTA walks one (batch, head, segment) in chunks of 128 tokens; two CTAs share an SM (256 TMEM columns and
// ~107 KB of shared memory each), so one CTA's tensor-core work overlaps the other's feature-map (MUFU) work.
// 4 worker warps (thread = TMEM lane = token row, or feature row of the state) + 1 control warp (TMA + MMA issue).
// Per chunk, three MMA batches, each committed to one mbarrier:
//   1. U_q = X_q Om, U_k = X_k Om                     SS, N = 64,  K = 64    -> TMEM [0,64) / [64,128)
//      workers: phi(k) -> smem [token][feature] (K-major B of 2, MN-major A of 3b); phi(q) -> TMEM [0,64) in place as
//      packed bf16 (the A operand of 2 and 3a); the normaliser's inter-chunk part phi(q).z in fp32 registers
//   2. S = phi(q) phi(k)^T                            TS, N = 128, K = 128   -> TMEM [64,192)
//      workers (while 2 runs): z += column sums of phi(k); then P = tril(S) -> bf16 in place (TMEM [64,128)), row sums
//   3a. O = phi(q) S'_bf16 + P V                      TS, N = 64,  K = 128 x 2 -> TMEM [128,192)
//   3b. S' += phi(k)^T V                              SS, N = 64,  K = 128   -> TMEM [192,256) (persistent state)
//      workers: out = O / den; S' -> bf16 -> smem (B operand of the next chunk's 3a)
// The normaliser column of the reference formulation (V' = [v | 1]) is carried as the fp32 vector z = sum phi(k) on
// the CUDA cores: den_i = rowsum(tril S)_i + phi(q_i).z_prev + 1e-6, which keeps every MMA at N = 64 / 128.
#include "tc_ptx.cuh"

namespace favor3 {
using namespace tcp;

constexpr int C = 128, FE = 64, FM = 128, FV = 80, NT = 160;
constexpr float F_EPS = 1e-6f, F_S2 = 0.125f, F_HALF_LOG_M = 2.4260151319598084f, K2 = 1.4426950408889634f;
// shared memory (offsets from the 1024-aligned base)
constexpr uint32_t OFF_XQ = 0, OFF_XK = 16384, OFF_XV = 32768, OFF_PK = 49152, OFF_SB = 81920, OFF_OM = 98304,
                   OFF_Z = 106496, OFF_ZP = OFF_Z + 512, OFF_BAR = OFF_ZP + 1024, SMEM_USED = OFF_BAR + 128;
constexpr int SMEM_BYTES = SMEM_USED + 1024;
static_assert(2 * (SMEM_BYTES + 1024) <= 227 * 1024, "two CTAs per SM");
// tensor memory columns
constexpr uint32_t T_PQ = 0, T_SC = 64, T_O = 128, T_ST = 192, T_COLS = 256;

struct Params {
  bf16* out;
  int64_t ld_out;
  float* den_out;
  const float* omega;
  const float* state_in;
  float* state_out;
  float* seg_states;
  int nseg, seg_chunks, T, H, items, omega_f16;
  long long* dbg_clk;
};

__global__ void __launch_bounds__(NT, 2)
favor_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps z / zpart accesses LDS / STS, not generic)
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(smem);
  const uint32_t sXQ = sb + OFF_XQ, sXK = sb + OFF_XK, sXV = sb + OFF_XV, sPK = sb + OFF_PK, sSB = sb + OFF_SB, sOM = sb + OFF_OM;
  float* z = reinterpret_cast<float*>(smem + OFF_Z);
  float* zpart = reinterpret_cast<float*>(smem + OFF_ZP);
  // One mbarrier PER MMA batch (each completes once per chunk).  With a single barrier for the three batches a worker
  // that is slow to poll for batch 3 of chunk n can be overtaken by the commit of batch 1 of chunk n + 1 (the control
  // warp only waits for the tensor pipe, not for the workers, in between): two phase flips later its parity test reads
  // "not yet", and it then waits for batch 2 -- which needs that very worker at [B1].  Seen as a watchdog trap.
  const uint32_t bar_qk = sb + OFF_BAR, bar_v = bar_qk + 8, bar_m1 = bar_qk + 16, bar_xf = bar_qk + 24, bar_m2 = bar_qk + 32,
                 bar_m3 = bar_qk + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 128) {
    prefetch_map(&tmQ); prefetch_map(&tmK); prefetch_map(&tmV);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_m1, 1); mbar_init(bar_m2, 1); mbar_init(bar_m3, 1); mbar_init(bar_xf, 128);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), T_COLS);
  // Omega * 64^(-1/4) * log2(e) -> bf16 [e][f] (f contiguous): the MN-major B operand of batch 1
  stage_omega(p.omega, sOM, tid, NT, p.omega_f16 != 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tl = tmem + ((uint32_t)(warp & 3) << 21);   // this warp's lane quadrant: (32 * warp) << 16

  const uint32_t ID_U = p.omega_f16 ? make_idesc(64, false, true, 128, true, false) : make_idesc(64, false, true);
  constexpr uint32_t ID_S = make_idesc(128, false, false),
                     ID_O = make_idesc(64, false, true), ID_ST = make_idesc(64, true, true);
  uint32_t ph_qk = 0, ph_v = 0, ph_m = 0, ph_xf = 0;     // parities of the next completion each role waits for (ph_m: per chunk, all three batch barriers)
  // control warp: the tensor-memory base as a warp-uniform value, operand descriptors built once
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
  const DescLH dXQ = make_desc_lh(sXQ, 0, 1024), dXK = make_desc_lh(sXK, 0, 1024), dOM = make_desc_lh(sOM, 8192, 1024),
               dPK0 = make_desc_lh(sPK, 0, 1024), dPK1 = make_desc_lh(sPK + 16384, 0, 1024), dSBmn = make_desc_lh(sSB, 16384, 1024),
               dXVmn = make_desc_lh(sXV, 16384, 1024), dPKmn = make_desc_lh(sPK, 16384, 1024);

  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int bh = item / p.nseg, seg = item % p.nseg;
    const int b = bh / p.H, h = bh % p.H;
    const int t_begin = seg * p.seg_chunks * C;
    const int t_end = (t_begin + p.seg_chunks * C < p.T) ? t_begin + p.seg_chunks * C : p.T;

    if (warp == 4) {
      if (t_begin < t_end && elect_one()) {
        mbar_expect_tx(bar_qk, 2 * C * FE * 2);
        tma_load_3d(&tmQ, bar_qk, sXQ, h * FE, t_begin, b);
        tma_load_3d(&tmK, bar_qk, sXK, h * FE, t_begin, b);
        mbar_expect_tx(bar_v, C * FE * 2);
        tma_load_3d(&tmV, bar_v, sXV, h * FE, t_begin, b);
      }
      __syncwarp();
    } else {
      // ---- prefix state of this item: fp32 -> TMEM, bf16 -> smem, z ----
      const float* sin = nullptr;
      if (p.state_in) sin = p.state_in + (int64_t)bh * FM * FV;
      else if (p.seg_states && seg > 0) sin = p.seg_states + ((int64_t)bh * (p.nseg + 1) + seg) * FM * FV;
      const float* row = sin ? sin + tid * FV : nullptr;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 t = row ? *reinterpret_cast<const float4*>(row + half * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          r[4 * j] = __float_as_uint(t.x); r[4 * j + 1] = __float_as_uint(t.y);
          r[4 * j + 2] = __float_as_uint(t.z); r[4 * j + 3] = __float_as_uint(t.w);
        }
        tmem_st32(tl + T_ST + half * 32, r);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 t;
          t.x = pack_bf16x2(__uint_as_float(r[8 * cc]), __uint_as_float(r[8 * cc + 1]));
          t.y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]), __uint_as_float(r[8 * cc + 3]));
          t.z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]), __uint_as_float(r[8 * cc + 5]));
          t.w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]), __uint_as_float(r[8 * cc + 7]));
          sts128(sSB + sw128(tid, half * 4 + cc), t);
        }
      }
      z[tid] = row ? row[FE] : 0.f;
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      named_bar_sync<2>(128);                  // z of every feature is in place before any row's phi(q).z
    }

    for (int t0 = t_begin; t0 < t_end; t0 += C) {
      const int valid = (p.T - t0 < C) ? (p.T - t0) : C;
      if (warp == 4) {
        // ============================ control warp: TMA + MMA issue ============================
        // all 32 lanes walk the protocol (waits, phase bits, descriptors: warp-uniform, so they live in uniform
        // registers); one elected lane issues -- see tc_ptx.cuh, elect_one
        mbar_wait(bar_qk, ph_qk); ph_qk ^= 1;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_PQ, desc_at(dXQ, ks * 32), desc_at(dOM, ks * 2048), ID_U, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_SC, desc_at(dXK, ks * 32), desc_at(dOM, ks * 2048), ID_U, ks > 0);
          umma_commit(bar_m1);
        }
        __syncwarp();
        // the next chunk's q, k once every worker has read its x rows (the row norms) and batch 1 is done with them
        mbar_wait(bar_xf, ph_xf); ph_xf ^= 1;
        if (t0 + C < t_end && elect_one()) {
          mbar_expect_tx(bar_qk, 2 * C * FE * 2);
          tma_load_3d(&tmQ, bar_qk, sXQ, h * FE, t0 + C, b);
          tma_load_3d(&tmK, bar_qk, sXK, h * FE, t0 + C, b);
        }
        __syncwarp();
        named_bar_sync<1>(NT);                 // [B1] phi(k) in smem, phi(q) in TMEM
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ts(tm + T_SC, tm + T_PQ + ks * 8, desc_at((ks >> 2) ? dPK1 : dPK0, (ks & 3) * 32), ID_S, ks > 0);
          umma_commit(bar_m2);
        }
        __syncwarp();
        named_bar_sync<1>(NT);                 // [B2] P in TMEM (and, from the previous chunk, S'_bf16 in smem)
        mbar_wait(bar_v, ph_v); ph_v ^= 1;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ts(tm + T_O, tm + T_PQ + ks * 8, desc_at(dSBmn, ks * 2048), ID_O, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ts(tm + T_O, tm + T_SC + ks * 8, desc_at(dXVmn, ks * 2048), ID_O, 1u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss(tm + T_ST, desc_at(dPKmn, ks * 2048), desc_at(dXVmn, ks * 2048), ID_ST, 1u);
          umma_commit(bar_m3);
        }
        __syncwarp();
        mbar_wait(bar_m3, ph_m);               // batch 3 done: v, phi(k) and the TMEM operands are free
        if (t0 + C < t_end && elect_one()) {
          mbar_expect_tx(bar_v, C * FE * 2);
          tma_load_3d(&tmV, bar_v, sXV, h * FE, t0 + C, b);
        }
        ph_m ^= 1;
        __syncwarp();
      } else {
        // ================================== worker warps ==================================
        const int i = tid;                     // token row of the chunk == TMEM lane
        const bool rowok = i < valid;
        int wk = 0;
// clock64 stamps of one worker thread (perf triage): compiled in only with -DEMO_KERNEL_DBG_CLK -- the run-time check alone
// was ~9 % of the attention backward's executed instructions (six stamps per tile in an issue-bound loop)
#ifdef EMO_KERNEL_DBG_CLK
#define FSTAMP() do { if (p.dbg_clk && tid == 0 && blockIdx.x == 0 && t0 == t_begin + 3 * C && wk < 16) p.dbg_clk[wk++] = clock64(); } while (0)
#else
#define FSTAMP() do { } while (0)
#endif
        FSTAMP();
        mbar_wait(bar_m1, ph_m);                     // batch 1: U_q, U_k (which also means q, k have landed)
        FSTAMP();
        tc_fence_after();
        float ssq = 0.f, ssk = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 a = lds128(sXQ + sw128(i, c)), bq = lds128(sXK + sw128(i, c));
          float f0, f1;
          unpack_bf16x2(a.x, f0, f1); ssq += f0 * f0 + f1 * f1; unpack_bf16x2(a.y, f0, f1); ssq += f0 * f0 + f1 * f1;
          unpack_bf16x2(a.z, f0, f1); ssq += f0 * f0 + f1 * f1; unpack_bf16x2(a.w, f0, f1); ssq += f0 * f0 + f1 * f1;
          unpack_bf16x2(bq.x, f0, f1); ssk += f0 * f0 + f1 * f1; unpack_bf16x2(bq.y, f0, f1); ssk += f0 * f0 + f1 * f1;
          unpack_bf16x2(bq.z, f0, f1); ssk += f0 * f0 + f1 * f1; unpack_bf16x2(bq.w, f0, f1); ssk += f0 * f0 + f1 * f1;
        }
        mbar_arrive_after_reads(bar_xf, ssq + ssk);
        const float oq = (0.5f * F_S2 * ssq + F_HALF_LOG_M) * K2, ok = (0.5f * F_S2 * ssk + F_HALF_LOG_M) * K2;
        FSTAMP();
        // ---- phi(k) -> smem (rows past the end of the sequence are zero: they must not enter the state) ----
        uint32_t uk[2][32];                     // both halves of U_k in flight at once: one TMEM round trip, not two
        tmem_ld32_issue(tl + T_SC, uk[0]);
        tmem_ld32_issue(tl + T_SC + 32, uk[1]);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t (&r)[32] = uk[half];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint4 tp, tm;
            uint32_t* pp = reinterpret_cast<uint32_t*>(&tp);
            uint32_t* pm = reinterpret_cast<uint32_t*>(&tm);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float u0 = __uint_as_float(r[8 * cc + 2 * e]), u1 = __uint_as_float(r[8 * cc + 2 * e + 1]);
              pp[e] = rowok ? pack_bf16x2(ex2(u0 - ok), ex2(u1 - ok)) : 0u;
              pm[e] = rowok ? pack_bf16x2(ex2(-u0 - ok), ex2(-u1 - ok)) : 0u;
            }
            sts128(sPK + sw128(i, half * 4 + cc), tp);
            sts128(sPK + 16384 + sw128(i, half * 4 + cc), tm);
          }
        }
        FSTAMP();
        // ---- phi(q) -> TMEM in place (both halves of U_q are read before anything is written) ----
        float dq = 0.f;
        {
          uint32_t r0[32], r1[32], pp[32], pm[32];
          tmem_ld32_issue(tl + T_PQ, r0);
          tmem_ld32_issue(tl + T_PQ + 32, r1);
          tmem_ld_wait();
          const float4* z4 = reinterpret_cast<const float4*>(z);
          float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {     // features 2j .. 2j + 3 (of each 64-feature half)
            const float4 zp0 = z4[j >> 1], zm0 = z4[16 + (j >> 1)], zp1 = z4[8 + (j >> 1)], zm1 = z4[24 + (j >> 1)];
            {
              const float u0 = __uint_as_float(r0[2 * j]), u1 = __uint_as_float(r0[2 * j + 1]);
              const float u2 = __uint_as_float(r0[2 * j + 2]), u3 = __uint_as_float(r0[2 * j + 3]);
              const float a0 = ex2(u0 - oq), a1 = ex2(u1 - oq), a2 = ex2(u2 - oq), a3 = ex2(u3 - oq);
              const float b0 = ex2(-u0 - oq), b1 = ex2(-u1 - oq), b2 = ex2(-u2 - oq), b3 = ex2(-u3 - oq);
              d0 += a0 * zp0.x + a1 * zp0.y; d1 += a2 * zp0.z + a3 * zp0.w;
              d2 += b0 * zm0.x + b1 * zm0.y; d3 += b2 * zm0.z + b3 * zm0.w;
              pp[j] = pack_bf16x2(a0, a1); pp[j + 1] = pack_bf16x2(a2, a3);
              pm[j] = pack_bf16x2(b0, b1); pm[j + 1] = pack_bf16x2(b2, b3);
            }
            {
              const float u0 = __uint_as_float(r1[2 * j]), u1 = __uint_as_float(r1[2 * j + 1]);
              const float u2 = __uint_as_float(r1[2 * j + 2]), u3 = __uint_as_float(r1[2 * j + 3]);
              const float a0 = ex2(u0 - oq), a1 = ex2(u1 - oq), a2 = ex2(u2 - oq), a3 = ex2(u3 - oq);
              const float b0 = ex2(-u0 - oq), b1 = ex2(-u1 - oq), b2 = ex2(-u2 - oq), b3 = ex2(-u3 - oq);
              d0 += a0 * zp1.x + a1 * zp1.y; d1 += a2 * zp1.z + a3 * zp1.w;
              d2 += b0 * zm1.x + b1 * zm1.y; d3 += b2 * zm1.z + b3 * zm1.w;
              pp[16 + j] = pack_bf16x2(a0, a1); pp[16 + j + 1] = pack_bf16x2(a2, a3);
              pm[16 + j] = pack_bf16x2(b0, b1); pm[16 + j + 1] = pack_bf16x2(b2, b3);
            }
          }
          dq = (d0 + d1) + (d2 + d3);
          tmem_st32(tl + T_PQ, pp);             // packed columns 0..31  = features 0..63   (exp(+u - o))
          tmem_st32(tl + T_PQ + 32, pm);        // packed columns 32..63 = features 64..127 (exp(-u - o))
        }
        FSTAMP();
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<1>(NT);                 // [B1]
        FSTAMP();
        // ---- z += column sums of phi(k), in the shadow of batch 2 ----
        {
          const int w = tid & 31, blk = (tid >> 5) & 1, hh = tid >> 6;
          const uint32_t base = sPK + blk * 16384 + (w & 3) * 4;
          const int cw = w >> 2;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
          for (int j = hh * 64; j < hh * 64 + 64; ++j) {
            float lo, hi;
            unpack_bf16x2(lds32(base + j * 128 + ((cw ^ (j & 7)) << 4)), lo, hi);
            s0 += lo; s1 += hi;
          }
          zpart[hh * 128 + blk * 64 + 2 * w] = s0;
          zpart[hh * 128 + blk * 64 + 2 * w + 1] = s1;
          named_bar_sync<2>(128);
          z[tid] += zpart[tid] + zpart[128 + tid];
        }
        FSTAMP();
        // ---- P = tril(S) -> bf16 in place; row sums ----
        mbar_wait(bar_m2, ph_m);                     // batch 2
        FSTAMP();
        tc_fence_after();
        float rs = 0.f;
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          uint32_t pk[16];
          if (pc <= warp) {                     // warp-uniform: key columns 32 pc .. are all in the future of rows < 32 pc
            uint32_t r[32];
            tmem_ld32_issue(tl + T_SC + pc * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c0 = pc * 32 + 2 * j;
              const float s0 = (c0 <= i) ? __uint_as_float(r[2 * j]) : 0.f;
              const float s1 = (c0 + 1 <= i) ? __uint_as_float(r[2 * j + 1]) : 0.f;
              rs += s0 + s1;
              pk[j] = pack_bf16x2(s0, s1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          tmem_st16(tl + T_SC + pc * 16, pk);
        }
        const float den = rs + dq + F_EPS;
        tmem_st_wait();
        tc_fence_before();
        FSTAMP();
        named_bar_sync<1>(NT);                 // [B2]
        FSTAMP();
        // ---- out = O / den; S' -> bf16 -> smem ----
        mbar_wait(bar_m3, ph_m); ph_m ^= 1;          // batch 3
        FSTAMP();
        tc_fence_after();
        {
          const float inv = 1.f / den;
          uint32_t ro[2][32];                   // both halves of O in flight at once
          tmem_ld32_issue(tl + T_O, ro[0]);
          tmem_ld32_issue(tl + T_O + 32, ro[1]);
          tmem_ld_wait();
          // The 32 output rows of this warp (128 bytes each) leave through the warp's OWN rows of the phi(k) tile (dead
          // since batch 3 completed; only this warp ever writes them): row per lane in, 8 lanes per row out, so that
          // a store instruction touches 4 lines instead of 32 (the load / store unit is shared by the SM's two CTAs).
          const uint32_t stg = sPK + (uint32_t)warp * 4096u;
          const int lane = tid & 31;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t (&r)[32] = ro[half];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint4 t;
              t.x = pack_bf16x2(__uint_as_float(r[8 * cc]) * inv, __uint_as_float(r[8 * cc + 1]) * inv);
              t.y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]) * inv, __uint_as_float(r[8 * cc + 3]) * inv);
              t.z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]) * inv, __uint_as_float(r[8 * cc + 5]) * inv);
              t.w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]) * inv, __uint_as_float(r[8 * cc + 7]) * inv);
              sts128(stg + sw128(lane, half * 4 + cc), t);
            }
          }
          __syncwarp();
          {
            const int sub = lane >> 3, ch = lane & 7, nv = valid - 32 * warp;
            bf16* obase = p.out + ((int64_t)b * p.T + t0 + 32 * warp) * p.ld_out + (int64_t)h * FE;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int row = 4 * j + sub;
              const uint4 t = lds128(stg + sw128(row, ch));
              if (row < nv) *(reinterpret_cast<uint4*>(obase + (int64_t)row * p.ld_out) + ch) = t;
            }
          }
          __syncwarp();
          if (p.den_out && rowok) p.den_out[((int64_t)b * p.T + t0 + i) * p.H + h] = den;
        }
        FSTAMP();
        const bool last = t0 + C >= t_end;
        float* so = nullptr;                    // fp32 copy of the final state: [128][80] = [S' | z | 0]
        float* so2 = nullptr;
        if (last && seg == p.nseg - 1) {
          if (p.state_out) so = p.state_out + (int64_t)bh * FM * FV;
          if (p.seg_states) so2 = p.seg_states + ((int64_t)bh * (p.nseg + 1) + p.nseg) * FM * FV;
        }
        uint32_t rs_[2][32];                    // both halves of the state row in flight at once
        tmem_ld32_issue(tl + T_ST, rs_[0]);
        tmem_ld32_issue(tl + T_ST + 32, rs_[1]);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t (&r)[32] = rs_[half];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint4 t;
            t.x = pack_bf16x2(__uint_as_float(r[8 * cc]), __uint_as_float(r[8 * cc + 1]));
            t.y = pack_bf16x2(__uint_as_float(r[8 * cc + 2]), __uint_as_float(r[8 * cc + 3]));
            t.z = pack_bf16x2(__uint_as_float(r[8 * cc + 4]), __uint_as_float(r[8 * cc + 5]));
            t.w = pack_bf16x2(__uint_as_float(r[8 * cc + 6]), __uint_as_float(r[8 * cc + 7]));
            sts128(sSB + sw128(tid, half * 4 + cc), t);
          }
          if (so || so2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                           __uint_as_float(r[4 * j + 3]));
              if (so) *reinterpret_cast<float4*>(so + tid * FV + half * 32 + 4 * j) = t;
              if (so2) *reinterpret_cast<float4*>(so2 + tid * FV + half * 32 + 4 * j) = t;
            }
          }
        }
        if (so || so2) {
          const float zf = z[tid];              // own element: updated by this thread before [B2]
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 t = make_float4(j == 0 ? zf : 0.f, 0.f, 0.f, 0.f);
            if (so) *reinterpret_cast<float4*>(so + tid * FV + FE + 4 * j) = t;
            if (so2) *reinterpret_cast<float4*>(so2 + tid * FV + FE + 4 * j) = t;
          }
        }
        fence_proxy_async();
        tc_fence_before();
        FSTAMP();
      }
    }
    if (warp < 4) {
      if (t_begin >= t_end && seg == p.nseg - 1) {          // empty last segment cannot happen (nseg is derived from T); keep the slots defined
      }
      if (p.seg_states && seg == 0) {                        // slot 0 = the empty prefix
        float4* s0 = reinterpret_cast<float4*>(p.seg_states + (int64_t)bh * (p.nseg + 1) * FM * FV);
        for (int i = tid; i < FM * FV / 4; i += 128) s0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      named_bar_sync<2>(128);                  // every worker is past its z / S'_bf16 reads before the next item rewrites them
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, T_COLS);
  }
}

}  // namespace favor3

// nseg segments of sc 128-token chunks per (b, h); the workspace / state arguments are those of emo_favor_fwd
int emo_favor_fwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, void* out,
                            int64_t ld_out, float* den, const float* state_in, float* state_out, float* seg_states,
                            int nseg, int sc, int B, int T_, int H, cudaStream_t s) {
  using namespace favor3;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_map_bt(&mq, q, (int64_t)H * FE, T_, B, ld, C))) return rc;
  if ((rc = make_map_bt(&mk, k, (int64_t)H * FE, T_, B, ld, C))) return rc;
  if ((rc = make_map_bt(&mv, v, (int64_t)H * FE, T_, B, ld, C))) return rc;
  Params p;
  p.out = (bf16*)out; p.ld_out = ld_out; p.den_out = den; p.omega = omega; p.state_in = state_in; p.state_out = state_out;
  p.seg_states = seg_states; p.nseg = nseg; p.seg_chunks = sc; p.T = T_; p.H = H; p.items = B * H * nseg; p.omega_f16 = favor_omega_f16();
  { const char* e = getenv("EMO_FAVOR_DBG_CLK"); p.dbg_clk = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
  const int max_ctas = 2 * emo_num_sms();
  const int grid = p.items < max_ctas ? p.items : max_ctas;
  favor_fwd_tc_kernel<<<grid, NT, SMEM_BYTES, s>>>(mq, mk, mv, p);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
