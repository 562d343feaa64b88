// K8 on the 5th-gen tensor cores (bf16, sm_100a): causal softmax attention of the stage-2 GPT-2 backbone
// (HF GPT2Attention._attn behind stage2_accompaniment/model/music_gpt2.py:42-51,84-87), flash style, every product a
// tcgen05.mma with its accumulator in tensor memory, q / k / v / dout tiles staged by TMA (128B swizzle).
// Same math, interface, dropout mask and lse convention as the mma.sync kernels in attn.cu, which stay as the fp32
// parity path, the relative-position (stage 1) path and the A/B switch EMO_ATTN_TC=0.
//
// forward   grid (q tiles, B*H), 2 CTAs per SM, 4 softmax warps (thread = query row = TMEM lane) + 1 control warp
//   per key tile of 128:  S = Q K^T (SS, N = 128)  ->  online softmax in registers, P (bf16, dropout applied) written
//   in place over S  ->  O += P V (TS: A operand from tensor memory).  O is rescaled lazily (only when a row maximum
//   grows by more than 2^8), so most tiles never touch it.  The second CTA of the SM fills the tensor pipe while this
//   one does its exponentials.
// backward  grid (k tiles, B*H), 1 CTA per SM, 8 worker warps (thread = key row, half of the query columns) + 1
//   control warp.  Per query tile, TRANSPOSED tiles so that every masked tile is an A operand without a transpose:
//     S^T = K Q^T, dP^T = V dO^T (SS, N = 128)  ->  P^T -> tensor memory, dS^T -> shared memory (bf16)
//     dV += P^T dO (TS)   dK += dS^T Q (SS, A K-major)   dQ_tile = dS K (SS, the same dS^T tile read as an MN-major A)
//   dK / dV accumulate in tensor memory over the CTA's query tiles; dQ tiles are added into an fp32 workspace with
//   vector reductions (red.global.add.v4.f32) and converted once at the end.  S^T / dP^T of the NEXT tile are issued as
//   soon as the workers hold the current tile in registers, so the tensor pipe runs under the element-wise pass.
#include "tc_ptx.cuh"

namespace attn3 {
using namespace tcp;

constexpr int BM = 128, BN = 128, HD = 64;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr uint32_t TILE = 16384;               // 128 rows x 64 bf16

// one 32-bit hash covers the aligned pair (e, e + 1), e even: low 16 bits decide e, high 16 bits e + 1 (common.cuh)
struct DropRow {
  uint32_t key, lo0;
  bool fast;
};
// e0 = flat index of the first element; fast path needs e0 even and no 32-bit wrap of the pair counter within n elements
__device__ __forceinline__ DropRow drop_row(uint64_t seed, uint64_t e0, int n) {
  DropRow d;
  const uint64_t pair0 = e0 >> 1;
  d.lo0 = (uint32_t)pair0;
  d.key = emo_drop_key(seed, (uint32_t)(pair0 >> 32));
  d.fast = ((e0 & 1) == 0) && (d.lo0 <= 0xFFFFFFFFu - (uint32_t)n);
  return d;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
namespace fwd {
constexpr int NT = 160;
constexpr uint32_t OFF_Q = 0, OFF_K = TILE, OFF_V = 3 * TILE, OFF_BAR = 5 * TILE, SMEM_USED = OFF_BAR + 128;
constexpr int SMEM_BYTES = SMEM_USED + 1024;
static_assert(2 * (SMEM_BYTES + 1024) <= 227 * 1024, "two CTAs per SM");
constexpr uint32_t T_S = 0, T_O = 128, T_COLS = 256;

struct Params {
  bf16* out; int64_t ld_o; float* lse;
  int Tq, Tk, H, nqt;
  float sl2, keep_scale;                       // scale * log2(e); 1 / (1 - p)
  uint32_t drop_thr; uint64_t seed;
  // stage-1 relative-position mode (relattn_tc.cu): bias [B*H][Tq][Tk] bf16 = the shifted position scores, added to
  // Q K^T before the scale; rel != 0 also switches dropout to "drop keys, then renormalise" (a softmax over the kept
  // keys: the mask joins the causal mask) and the 1e-8 of optimus_txl_decoder.py:362-363
  const bf16* bias;
  int rel;
};

// REL / DROP (0 = none, 1 = one hash per aligned key pair, 2 = per-element hashes: odd Tk or > 2^33 scores) are
// compile-time: with every variant in one body the kernel sat at the 168-register cap with 216 bytes of spills
template <bool REL, int DROP>
__global__ void __launch_bounds__(NT, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(smem);
  const uint32_t sQ = sb + OFF_Q, sK = sb + OFF_K, sV = sb + OFF_V;
  const uint32_t bar_q = sb + OFF_BAR, bar_k = bar_q + 8 /* x2 */, bar_v = bar_q + 24 /* x2 */, bar_s = bar_q + 40, bar_o = bar_q + 48,
                 bar_p = bar_q + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = p.nqt - 1 - (int)blockIdx.x;                  // long rows first
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int i0 = qt * BM, off = p.Tk - p.Tq;
  const int kmax = (p.Tk < i0 + BM + off) ? p.Tk : i0 + BM + off;     // keys [0, kmax) are visible to some row of the tile
  const int n = (kmax + BN - 1) / BN;

  if (tid == 128) {
    prefetch_map(&tmQ); prefetch_map(&tmK); prefetch_map(&tmV);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_k + 8, 1); mbar_init(bar_v, 1); mbar_init(bar_v + 8, 1);
    mbar_init(bar_s, 1); mbar_init(bar_o, 1); mbar_init(bar_p, 128);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), T_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t ID_S = make_idesc(128, false, false), ID_O = make_idesc(64, false, true);

  if (warp == 4) {
    // ================================ control warp: all lanes walk the loop, one elected lane issues ================================
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const DescLH dQ = make_desc_lh(sQ, 0, 1024);
    if (elect_one()) {
      mbar_expect_tx(bar_q, TILE);
      tma_load_3d(&tmQ, bar_q, sQ, h * HD, i0, b);
      mbar_expect_tx(bar_k, TILE);
      tma_load_3d(&tmK, bar_k, sK, h * HD, 0, b);
      mbar_expect_tx(bar_v, TILE);
      tma_load_3d(&tmV, bar_v, sV, h * HD, 0, b);
    }
    __syncwarp();
    mbar_wait(bar_q, 0);
    for (int it = 0; it < n; ++it) {
      const uint32_t st = it & 1, par = (it >> 1) & 1;
      const DescLH dK = make_desc_lh(sK + st * TILE, 0, 1024), dV = make_desc_lh(sV + st * TILE, 16384, 1024);
      mbar_wait(bar_k + 8 * st, par);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_S, desc_at(dQ, ks * 32), desc_at(dK, ks * 32), ID_S, ks > 0);
        umma_commit(bar_s);
        if (it + 1 < n) {                       // S(it - 1), the last reader of that stage, was seen complete by the workers
          mbar_expect_tx(bar_k + 8 * (st ^ 1), TILE);
          tma_load_3d(&tmK, bar_k + 8 * (st ^ 1), sK + (st ^ 1) * TILE, h * HD, (it + 1) * BN, b);
        }
      }
      __syncwarp();
      mbar_wait(bar_p, it & 1);                 // P(it) is in tensor memory (and O rescaled if it had to be)
      if (it + 1 < n && elect_one()) {          // P V(it - 1) completed before S(it) did
        mbar_expect_tx(bar_v + 8 * (st ^ 1), TILE);
        tma_load_3d(&tmV, bar_v + 8 * (st ^ 1), sV + (st ^ 1) * TILE, h * HD, (it + 1) * BN, b);
      }
      __syncwarp();
      mbar_wait(bar_v + 8 * st, par);
      tc_fence_after();
      const uint32_t acc = it > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_ts(tm + T_O, tm + T_S + ks * 8, desc_at(dV, ks * 2048), ID_O, (ks > 0) ? 1u : acc);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax warps: thread = query row ================================
    const int i = i0 + tid;
    const uint32_t tl = tmem + ((uint32_t)warp << 21);
    float m_used = -INFINITY, l = 0.f;
    const int lim_base = i + off;               // key j is visible iff j <= lim_base
    for (int it = 0; it < n; ++it) {
      const int j0 = it * BN;
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32_issue(tl + T_S + 32 * c, s + 32 * c);
      tmem_ld_wait();
      if (REL && p.bias) {                      // + position scores of this row, 32 columns at a time
        // blocked plane (relattn_tc.cu, rel_blocked_off): chunk c of the warp's 32 rows is 512 contiguous bytes, so a load
        // instruction touches 4 lines (row-major rows made it 32: the LSU, shared by the SM's two CTAs, was the bound)
        const int nct = (p.Tk + BN - 1) / BN;
        const uint4* bblk = reinterpret_cast<const uint4*>(p.bias + (((int64_t)bh * p.nqt + qt) * nct + it) * (int64_t)(BM * BN)) +
                            (warp * 16) * 32 + lane;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4 bb[4];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) bb[k4] = __ldg(bblk + (4 * c4 + k4) * 32);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t w[4] = {bb[k4].x, bb[k4].y, bb[k4].z, bb[k4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float b0, b1;
              unpack_bf16x2(w[e], b0, b1);
              const int c = 32 * c4 + 8 * k4 + 2 * e;
              s[c] = __float_as_uint(__uint_as_float(s[c]) + b0);
              s[c + 1] = __float_as_uint(__uint_as_float(s[c + 1]) + b1);
            }
          }
        }
      }
      if (REL && DROP) {                        // dropped keys leave the softmax
        const uint64_t e0 = ((uint64_t)bh * p.Tq + i) * (uint64_t)p.Tk + j0;
        const DropRow dr = drop_row(p.seed, e0, BN);
        if (DROP == 1) {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const uint32_t hsh = emo_drop_mix(dr.lo0 + c, dr.key);
            if ((hsh & 0xffffu) < p.drop_thr) s[2 * c] = 0xff800000u;
            if ((hsh >> 16) < p.drop_thr) s[2 * c + 1] = 0xff800000u;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (!emo_drop_keep(p.seed, e0 + c, p.drop_thr)) s[c] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
      if (j0 + BN - 1 > i0 + off) {             // CTA-uniform: the tile crosses the diagonal
        const int lim = lim_base - j0;
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const float v = (c <= lim) ? __uint_as_float(s[c]) : -INFINITY;
          s[c] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
      }
      const float m_tile = mx * p.sl2;
      const bool need = m_tile > m_used + 8.f;  // lazy rescale: P stays below 2^8, the row sum absorbs the stale maximum
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? m_tile : m_used;
        const float alpha = ex2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
        if (it > 0) {                           // P V(it - 1) is complete (S(it) was committed after it)
          // 16 columns at a time: the 128 scores of the row are live here, and a 64-register copy of O on top of them
          // spilled (216 bytes of local memory in the hot loop); the rescale itself is rare (lazy, above)
#pragma unroll 1
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t o[16];
            tmem_ld16_issue(tl + T_O + 16 * q4, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st16(tl + T_O + 16 * q4, o);
          }
        }
      }
      const float nm = (m_used == -INFINITY) ? 0.f : -m_used;     // every key so far masked / dropped: exp2(-inf + 0) = 0, not NaN
      float sum0 = 0.f, sum1 = 0.f;
      const uint64_t e0 = ((uint64_t)bh * p.Tq + i) * (uint64_t)p.Tk + j0;
      const DropRow dr = drop_row(p.seed, e0, BN);
      (void)dr;
      // P leaves for tensor memory 32 packed words at a time: the scores die as they are consumed, so the row never
      // needs more than its 128 score registers
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[32];
#pragma unroll
        for (int cl = 0; cl < 32; ++cl) {
          const int c = 32 * hf + cl;
          const float p0 = ex2(fmaf(__uint_as_float(s[2 * c]), p.sl2, nm)), p1 = ex2(fmaf(__uint_as_float(s[2 * c + 1]), p.sl2, nm));
          sum0 += p0; sum1 += p1;
          if (DROP == 0 || REL) {
            pk[cl] = pack_bf16x2(p0, p1);
          } else if (DROP == 1) {
            const uint32_t hsh = emo_drop_mix(dr.lo0 + c, dr.key);
            pk[cl] = pack_bf16x2(((hsh & 0xffffu) >= p.drop_thr) ? p0 * p.keep_scale : 0.f, ((hsh >> 16) >= p.drop_thr) ? p1 * p.keep_scale : 0.f);
          } else {
            pk[cl] = pack_bf16x2(emo_drop_keep(p.seed, e0 + 2 * c, p.drop_thr) ? p0 * p.keep_scale : 0.f,
                                 emo_drop_keep(p.seed, e0 + 2 * c + 1, p.drop_thr) ? p1 * p.keep_scale : 0.f);
          }
        }
        tmem_st32(tl + T_S + 32 * hf, pk);
      }
      l += sum0 + sum1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- out = O / l, lse ----
    mbar_wait(bar_o, (n - 1) & 1);
    tc_fence_after();
    {
      // stage 1: P / (sum P + 1e-8) with sum P = 1, or 0 when every key of the row was dropped
      const float inv = REL ? (l > 0.f ? 1.f / (l * (1.f + 1e-8f)) : 0.f) : 1.f / l;
      bf16* orow = p.out + ((int64_t)b * p.Tq + i) * p.ld_o + (int64_t)h * HD;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        tmem_ld32_issue(tl + T_O + half * 32, o);
        tmem_ld_wait();
        if (i < p.Tq) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint4 t;
            t.x = pack_bf16x2(__uint_as_float(o[8 * cc]) * inv, __uint_as_float(o[8 * cc + 1]) * inv);
            t.y = pack_bf16x2(__uint_as_float(o[8 * cc + 2]) * inv, __uint_as_float(o[8 * cc + 3]) * inv);
            t.z = pack_bf16x2(__uint_as_float(o[8 * cc + 4]) * inv, __uint_as_float(o[8 * cc + 5]) * inv);
            t.w = pack_bf16x2(__uint_as_float(o[8 * cc + 6]) * inv, __uint_as_float(o[8 * cc + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + half * 32 + cc * 8) = t;
          }
        }
      }
      if (p.lse && i < p.Tq) p.lse[(int64_t)bh * p.Tq + i] = l > 0.f ? m_used * LN2 + logf(l) : INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, T_COLS);
  }
}
}  // namespace fwd

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
namespace bwd {
constexpr int NW = 16, NT = 32 * (NW + 1), NST = 3;      // 16 worker warps (4 per TMEM lane quadrant) + 1 control warp
constexpr uint32_t OFF_K = 0, OFF_V = TILE, OFF_Q = 2 * TILE /* x3 */, OFF_DO = 5 * TILE /* x3 */, OFF_DS = 8 * TILE /* x2 */,
                   OFF_DQ = 10 * TILE /* 16 warps x [32 rows][16 fp32] */, OFF_L = 12 * TILE /* lse2[3][128] */, OFF_D = OFF_L + NST * 512,
                   OFF_BAR = OFF_D + NST * 512, SMEM_USED = OFF_BAR + 128;
constexpr int SMEM_BYTES = SMEM_USED + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");
constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 320, T_DQ = 384, T_P = 448, T_COLS = 512;
constexpr uint32_t LD_BYTES = 2 * TILE + 2 * 512;

struct Params {
  const float* lse2; const float* dsum;        // [B*H][Tq_pad]: lse * log2(e) (+inf past Tq), scale * rowsum(dO * O)
  bf16* dk; bf16* dv; int64_t ld_dkv;
  int Tq, Tk, Tq_pad, H, nqt;
  float scale, sl2, keep_scale;
  uint32_t drop_thr; uint64_t seed;
  int drop_fast;                               // Tk even and fewer than 2^32 element pairs: one 32-bit hash per (query, key pair)
  // stage-1 relative-position mode (relattn_tc.cu): biasT [B*H][Tk][Tq] bf16 = the TRANSPOSED shifted position scores,
  // added to K Q^T; dbiasT (same layout) receives dS^T; rel != 0: dropout = drop keys and renormalise (no 1/(1-p))
  const bf16* biasT; bf16* dbiasT;
  int rel;
  long long* dbg_clk;                          // optional: clock64 stamps of worker thread 0, tiles 3 and 4 of CTA (0, 0)
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Element-wise pass of one thread: 32 query columns of key row j.  s = raw S^T, dp = raw dP^T (fp32 bits);
// P = exp2(s * sl2 - lse2), dS = P * (dP_dropped - D) * scale  (aD holds D * scale).  DROP == 1: the fast mask (the two
// lanes of a key pair share one hash per query column: each computes every other column and they swap);
// DROP == 2: the general per-element mask.
template <bool DIAG, int DROP, bool REL = false>
__device__ __forceinline__ void ew_pass(const uint32_t (&s)[32], const uint32_t (&dp)[32], uint32_t aL, uint32_t aD, int cmin,
                                        const Params& p, uint32_t pbase, uint32_t key, int lane, uint64_t e_base,
                                        uint32_t (&pk)[16], uint32_t (&dsk)[16]) {
  const float kss = p.keep_scale * p.scale;
  const uint32_t odd = lane & 1, sel = odd ? 0x4432u : 0x4410u, pstep = (uint32_t)p.Tk >> 1;
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {                           // 4 query columns at a time
    const uint4 l4 = lds128(aL + 16 * c4), d4 = lds128(aD + 16 * c4);
    const float lv[4] = {__uint_as_float(l4.x), __uint_as_float(l4.y), __uint_as_float(l4.z), __uint_as_float(l4.w)};
    const float dv_[4] = {__uint_as_float(d4.x), __uint_as_float(d4.y), __uint_as_float(d4.z), __uint_as_float(d4.w)};
    bool keep[4] = {true, true, true, true};
    if (DROP == 1) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint32_t hm = emo_drop_mix(pbase + (uint32_t)(4 * c4 + 2 * m + odd) * pstep, key);
        const uint32_t ho = __shfl_xor_sync(0xffffffffu, hm, 1);
        keep[2 * m] = __byte_perm(odd ? ho : hm, 0, sel) >= p.drop_thr;
        keep[2 * m + 1] = __byte_perm(odd ? hm : ho, 0, sel) >= p.drop_thr;
      }
    } else if (DROP == 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) keep[e] = emo_drop_keep(p.seed, e_base + (uint64_t)(4 * c4 + e) * (uint64_t)p.Tk, p.drop_thr);
    }
    float pd[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * c4 + e;
      float pv = ex2(fmaf(__uint_as_float(s[c]), p.sl2, -lv[e]));
      if (DIAG) pv = (c >= cmin) ? pv : 0.f;
      const float dpv = __uint_as_float(dp[c]);
      if (DROP && REL) {                     // a dropped key is outside the softmax: P = 0 there
        pv = keep[e] ? pv : 0.f;
        pd[e] = pv;
        ds[e] = pv * fmaf(dpv, p.scale, -dv_[e]);
      } else if (DROP) {
        pd[e] = keep[e] ? pv * p.keep_scale : 0.f;
        ds[e] = pv * (keep[e] ? fmaf(dpv, kss, -dv_[e]) : -dv_[e]);
      } else {
        pd[e] = pv;
        ds[e] = pv * fmaf(dpv, p.scale, -dv_[e]);
      }
    }
    pk[2 * c4] = pack_bf16x2(pd[0], pd[1]); pk[2 * c4 + 1] = pack_bf16x2(pd[2], pd[3]);
    dsk[2 * c4] = pack_bf16x2(ds[0], ds[1]); dsk[2 * c4 + 1] = pack_bf16x2(ds[2], ds[3]);
  }
}

__global__ void __launch_bounds__(NT, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(smem);
  const uint32_t sK = sb + OFF_K, sV = sb + OFF_V, sQ = sb + OFF_Q, sDO = sb + OFF_DO, sDS = sb + OFF_DS, sDQ = sb + OFF_DQ, sL = sb + OFF_L,
                 sD = sb + OFF_D;
  const uint32_t bar_kv = sb + OFF_BAR, bar_ld = bar_kv + 8 /* x3 */, bar_s = bar_kv + 32, bar_sfree = bar_kv + 40, bar_p = bar_kv + 48,
                 bar_acc = bar_kv + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x;                                   // key tile 0 sees the most query tiles: long CTAs first
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int j0 = kt * BN, off = p.Tk - p.Tq;
  const int qt0 = (j0 - off > 0) ? (j0 - off) / BM : 0;        // first query tile with a row that sees key j0
  const int n = p.nqt - qt0;                                   // >= 1 (Tk >= Tq)

  if (tid == 32 * NW) {
    prefetch_map(&tmQ); prefetch_map(&tmK); prefetch_map(&tmV); prefetch_map(&tmDO); prefetch_map(&tmDQ);
    mbar_init(bar_kv, 1);
    for (int s_ = 0; s_ < NST; ++s_) mbar_init(bar_ld + 8 * s_, 1);
    mbar_init(bar_s, 1); mbar_init(bar_sfree, 32 * NW); mbar_init(bar_p, 32 * NW); mbar_init(bar_acc, 1);
    mbar_init_fence();
  }
  if (warp == NW) tmem_alloc(smem_u32(tmem_slot), T_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t ID_T = make_idesc(128, false, false),     // S^T, dP^T: both operands K-major
                     ID_KM = make_idesc(64, false, true),      // dV, dK: A K-major (TMEM / smem), B MN-major
                     ID_MM = make_idesc(64, true, true);       // dQ: A = dS^T tile read MN-major, B MN-major

  if (warp == NW) {
    // ================================ control warp: all lanes walk the loop, one elected lane issues ================================
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const DescLH dK_k = make_desc_lh(sK, 0, 1024), dV_k = make_desc_lh(sV, 0, 1024), dK_mn = make_desc_lh(sK, 16384, 1024),
                 dDS_k0 = make_desc_lh(sDS, 0, 1024), dDS_k1 = make_desc_lh(sDS + TILE, 0, 1024), dDS_mn = make_desc_lh(sDS, 16384, 1024);
    auto load_tile = [&](int t) {
      const uint32_t st = t % NST, bar = bar_ld + 8 * st;
      const int i0 = (qt0 + t) * BM;
      if (elect_one()) {
        mbar_expect_tx(bar, LD_BYTES);
        tma_load_3d(&tmQ, bar, sQ + st * TILE, h * HD, i0, b);
        tma_load_3d(&tmDO, bar, sDO + st * TILE, h * HD, i0, b);
        bulk_load(sL + st * 512, p.lse2 + (int64_t)bh * p.Tq_pad + i0, 512, bar);
        bulk_load(sD + st * 512, p.dsum + (int64_t)bh * p.Tq_pad + i0, 512, bar);
      }
      __syncwarp();
    };
    auto issue_scores = [&](int t) {
      const uint32_t st = t % NST;
      const DescLH dQ_k = make_desc_lh(sQ + st * TILE, 0, 1024), dDO_k = make_desc_lh(sDO + st * TILE, 0, 1024);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_S, desc_at(dK_k, ks * 32), desc_at(dQ_k, ks * 32), ID_T, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ss(tm + T_DP, desc_at(dV_k, ks * 32), desc_at(dDO_k, ks * 32), ID_T, ks > 0);
        umma_commit(bar_s);
      }
      __syncwarp();
    };
    if (elect_one()) {
      mbar_expect_tx(bar_kv, 2 * TILE);
      tma_load_3d(&tmK, bar_kv, sK, h * HD, j0, b);
      tma_load_3d(&tmV, bar_kv, sV, h * HD, j0, b);
    }
    __syncwarp();
    load_tile(0);
    if (n > 1) load_tile(1);
    mbar_wait(bar_kv, 0);
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    issue_scores(0);
    for (int it = 0; it < n; ++it) {
      const uint32_t st = it % NST;
      if (it + 1 < n) {                         // the workers hold S^T / dP^T (it) in registers: the next pair may overwrite them
        mbar_wait(bar_sfree, it & 1);
        mbar_wait(bar_ld + 8 * ((it + 1) % NST), ((it + 1) / NST) & 1);
        tc_fence_after();
        issue_scores(it + 1);
      }
      mbar_wait(bar_p, it & 1);                 // P^T (it) in tensor memory, dS^T (it) in shared memory, dQ (it - 1) read out
      if (it >= 1 && it + 2 < n) mbar_wait(bar_acc, (it - 1) & 1);    // stage of tile it - 1 is free (waited before this tile's commit)
      tc_fence_after();
      const uint32_t acc = it > 0 ? 1u : 0u;
      const DescLH dDO_mn = make_desc_lh(sDO + st * TILE, 16384, 1024), dQ_mn = make_desc_lh(sQ + st * TILE, 16384, 1024);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_ts(tm + T_DV, tm + T_P + ks * 8, desc_at(dDO_mn, ks * 2048), ID_KM, (ks > 0) ? 1u : acc);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss(tm + T_DK, desc_at((ks >> 2) ? dDS_k1 : dDS_k0, (ks & 3) * 32), desc_at(dQ_mn, ks * 2048), ID_KM, (ks > 0) ? 1u : acc);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_ss(tm + T_DQ, desc_at(dDS_mn, ks * 2048), desc_at(dK_mn, ks * 2048), ID_MM, ks > 0);
        umma_commit(bar_acc);
      }
      __syncwarp();
      if (it + 2 < n) load_tile(it + 2);
    }
  } else {
    // ================================ workers: thread = (key row r, column quarter ch) ================================
    const int quad = warp & 3, ch = warp >> 2;
    const int r = quad * 32 + lane;
    const int j = j0 + r;
    const uint32_t tl = tmem + ((uint32_t)quad << 21);
    const uint32_t sDQw = sDQ + warp * 2048;                 // this warp's [32 queries][16 head columns] fp32 staging tile (64B swizzle)
    const uint32_t key = emo_drop_key(p.seed, 0);
    const int drop_mode = p.drop_thr ? (p.drop_fast ? 1 : 2) : 0;
    // dQ tile (lanes = queries, 16 of the 64 head columns) += into the fp32 workspace: TMEM -> smem -> TMA reduce
    auto dq_out = [&](int t) {
      if (lane == 0) tma_store_wait_read<0>();               // the previous reduction has read the staging tile
      __syncwarp();
      uint32_t d[16];
      tmem_ld16_issue(tl + T_DQ + 16 * ch, d);
      tmem_ld_wait();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint4 t4;
        t4.x = d[4 * cc]; t4.y = d[4 * cc + 1]; t4.z = d[4 * cc + 2]; t4.w = d[4 * cc + 3];
        sts128(sDQw + (uint32_t)(lane * 64 + ((cc ^ ((lane >> 1) & 3)) << 4)), t4);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_reduce_add_3d(&tmDQ, sDQw, h * HD + 16 * ch, (qt0 + t) * BM + 32 * quad, b);
    };
    int wk = 0;
// clock64 stamps of one worker thread (perf triage): compiled in only with -DEMO_KERNEL_DBG_CLK -- the run-time check alone
// was ~9 % of the attention backward's executed instructions (six stamps per tile in an issue-bound loop)
#ifdef EMO_KERNEL_DBG_CLK
#define WSTAMP() do { if (p.dbg_clk && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && (it == 3 || it == 4) && wk < 16) p.dbg_clk[16 + wk++] = clock64(); } while (0)
#else
#define WSTAMP() do { } while (0)
#endif
    for (int it = 0; it < n; ++it) {
      const uint32_t st = it % NST;
      const int i0 = (qt0 + it) * BM;
      WSTAMP();
      mbar_wait(bar_s, it & 1);
      WSTAMP();
      tc_fence_after();
      mbar_wait(bar_ld + 8 * st, (it / NST) & 1);          // lse2 / dsum of this tile (TMA writes) are visible to this thread
      uint32_t s[32], dp[32];
      tmem_ld32_issue(tl + T_S + 32 * ch, s);
      tmem_ld32_issue(tl + T_DP + 32 * ch, dp);
      tmem_ld_wait();
      WSTAMP();
      tc_fence_before();
      mbar_arrive(bar_sfree);
      const uint32_t aL = sL + st * 512 + ch * 128, aD = sD + st * 512 + ch * 128;
      const bool diag = i0 + off < j0 + BN - 1;              // CTA-uniform: some (query, key) pair of the tile is masked
      const int cmin = j - off - i0 - 32 * ch;               // column c (query i0 + 32 ch + c) sees key j iff c >= cmin
      const uint64_t e_base = ((uint64_t)bh * p.Tq + (uint64_t)(i0 + 32 * ch)) * (uint64_t)p.Tk + (uint64_t)j;   // mask index of column 0
      const uint32_t pbase = (uint32_t)(e_base >> 1);
      // biasT / dbiasT: the 32 queries [i0 + 32 ch, + 32) of key row j, as chunks 4 ch .. 4 ch + 3 of row r of block (key tile,
      // query tile) of the blocked plane (relattn_tc.cu): 512 contiguous bytes per warp and chunk
      const int nqt_all = (p.Tq + BM - 1) / BM;
      const int64_t boff = (((int64_t)bh * ((p.Tk + BN - 1) / BN) + j0 / BN) * nqt_all + (qt0 + it)) * (int64_t)(BM * BN) +
                           ((int64_t)((quad * 16 + 4 * ch) * 32 + lane) << 3);
      if (p.biasT) {
        const uint4* brow = reinterpret_cast<const uint4*>(p.biasT + boff);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint4 bb = __ldg(brow + k4 * 32);
          const uint32_t w[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float b0, b1;
            unpack_bf16x2(w[e], b0, b1);
            s[8 * k4 + 2 * e] = __float_as_uint(__uint_as_float(s[8 * k4 + 2 * e]) + b0);
            s[8 * k4 + 2 * e + 1] = __float_as_uint(__uint_as_float(s[8 * k4 + 2 * e + 1]) + b1);
          }
        }
      }
      uint32_t pk[16], dsk[16];
      if (p.rel) {                                           // stage 1 (short sequences): one variant per dropout mode
        if (drop_mode == 0) ew_pass<true, 0, true>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
        else if (drop_mode == 1) ew_pass<true, 1, true>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
        else ew_pass<true, 2, true>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
      } else if (drop_mode == 0) {
        if (diag) ew_pass<true, 0>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
        else ew_pass<false, 0>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
      } else if (drop_mode == 1) {
        if (diag) ew_pass<true, 1>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
        else ew_pass<false, 1>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
      } else {
        ew_pass<true, 2>(s, dp, aL, aD, cmin, p, pbase, key, lane, e_base, pk, dsk);
      }
      if (p.dbiasT && j < p.Tk && i0 + 32 * ch < p.Tq) {     // d(position scores)^T = dS^T: 64 contiguous bytes per thread (Tq % 32 == 0)
        uint4* drow = reinterpret_cast<uint4*>(p.dbiasT + boff);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) drow[k4 * 32] = make_uint4(dsk[4 * k4], dsk[4 * k4 + 1], dsk[4 * k4 + 2], dsk[4 * k4 + 3]);
      }
      if (it > 0) {                                          // products of tile it - 1 are complete: P^T / dS^T / dQ may be reused
        mbar_wait(bar_acc, (it - 1) & 1);
        WSTAMP();
        tc_fence_after();
        dq_out(it - 1);
      }
      WSTAMP();
      tmem_st16(tl + T_P + 16 * ch, pk);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint4 t;
        t.x = dsk[4 * cc]; t.y = dsk[4 * cc + 1]; t.z = dsk[4 * cc + 2]; t.w = dsk[4 * cc + 3];
        sts128(sDS + (ch >> 1) * TILE + sw128(r, (ch & 1) * 4 + cc), t);
      }
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_p);
      WSTAMP();
    }
    mbar_wait(bar_acc, (n - 1) & 1);
    tc_fence_after();
    dq_out(n - 1);
    // ---- dK, dV rows of this key tile ----
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t d[16];
      tmem_ld16_issue(tl + (which ? T_DK : T_DV) + 16 * ch, d);
      tmem_ld_wait();
      if (j < p.Tk) {
        bf16* dst = (which ? p.dk : p.dv) + ((int64_t)b * p.Tk + j) * p.ld_dkv + (int64_t)h * HD + 16 * ch;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint4 t;
          t.x = pack_bf16x2(__uint_as_float(d[8 * cc]), __uint_as_float(d[8 * cc + 1]));
          t.y = pack_bf16x2(__uint_as_float(d[8 * cc + 2]), __uint_as_float(d[8 * cc + 3]));
          t.z = pack_bf16x2(__uint_as_float(d[8 * cc + 4]), __uint_as_float(d[8 * cc + 5]));
          t.w = pack_bf16x2(__uint_as_float(d[8 * cc + 6]), __uint_as_float(d[8 * cc + 7]));
          *reinterpret_cast<uint4*>(dst + cc * 8) = t;
        }
      }
    }
    if (lane == 0) tma_store_wait_all();                     // the last reduction has left shared memory before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) {
    tc_fence_after();
    tmem_dealloc(tmem, T_COLS);
  }
}

// lse2 = lse * log2(e) (+inf on the padding rows: P = 0 there), dsum = scale * rowsum(dO * O); 8 threads per row
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, int64_t ld_o,
                                     const float* __restrict__ lse, float* __restrict__ lse2, float* __restrict__ dsum,
                                     int Tq, int Tq_pad, int H, int64_t rows_pad, float scale) {
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;      // 8 threads per row
  const int part = threadIdx.x & 7;
  if (gw >= rows_pad) return;
  const int64_t bh = gw / Tq_pad;
  const int i = (int)(gw % Tq_pad);
  const int b = (int)(bh / H), h = (int)(bh % H);
  float dot = 0.f;
  if (i < Tq) {
    const int64_t o = ((int64_t)b * Tq + i) * ld_o + (int64_t)h * HD + part * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(out + o), d = *reinterpret_cast<const uint4*>(dout + o);
    float a0, a1, d0, d1;
    unpack_bf16x2(a.x, a0, a1); unpack_bf16x2(d.x, d0, d1); dot += a0 * d0 + a1 * d1;
    unpack_bf16x2(a.y, a0, a1); unpack_bf16x2(d.y, d0, d1); dot += a0 * d0 + a1 * d1;
    unpack_bf16x2(a.z, a0, a1); unpack_bf16x2(d.z, d0, d1); dot += a0 * d0 + a1 * d1;
    unpack_bf16x2(a.w, a0, a1); unpack_bf16x2(d.w, d0, d1); dot += a0 * d0 + a1 * d1;
  }
  dot += __shfl_xor_sync(0xffffffffu, dot, 1);
  dot += __shfl_xor_sync(0xffffffffu, dot, 2);
  dot += __shfl_xor_sync(0xffffffffu, dot, 4);
  if (part == 0) {
    dsum[gw] = dot * scale;
    lse2[gw] = (i < Tq) ? lse[bh * Tq + i] * LOG2E : INFINITY;
  }
}

// fp32 dQ workspace [B][Tq][H*64] (+ optional bf16 addend of the same dense shape) -> bf16 dq (row stride ld_dq)
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, const bf16* __restrict__ add, bf16* __restrict__ dq, int64_t ld_dq,
                                           int64_t rows, int HD_all) {
  const int vpr = HD_all / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * vpr) return;
  const int64_t row = idx / vpr;
  const int c = (int)(idx % vpr) * 8;
  float4 a = *reinterpret_cast<const float4*>(acc + row * HD_all + c), b = *reinterpret_cast<const float4*>(acc + row * HD_all + c + 4);
  if (add) {
    const uint4 t = *reinterpret_cast<const uint4*>(add + row * HD_all + c);
    float x0, x1;
    unpack_bf16x2(t.x, x0, x1); a.x += x0; a.y += x1;
    unpack_bf16x2(t.y, x0, x1); a.z += x0; a.w += x1;
    unpack_bf16x2(t.z, x0, x1); b.x += x0; b.y += x1;
    unpack_bf16x2(t.w, x0, x1); b.z += x0; b.w += x1;
  }
  uint4 t;
  t.x = pack_bf16x2(a.x, a.y); t.y = pack_bf16x2(a.z, a.w); t.z = pack_bf16x2(b.x, b.y); t.w = pack_bf16x2(b.z, b.w);
  *reinterpret_cast<uint4*>(dq + row * ld_dq + c) = t;
}
}  // namespace bwd
}  // namespace attn3

// ------------------------------------------------------------------------------------------------------------------
// host side (called from attn.cu's C ABI entry points)
// ------------------------------------------------------------------------------------------------------------------
static int g_attn_tc = -1;
int emo_attn_tc_enabled() {
  if (g_attn_tc < 0) { const char* e = getenv("EMO_ATTN_TC"); g_attn_tc = e ? atoi(e) : 1; }
  return g_attn_tc;
}
extern "C" void emo_attn_set_tc(int on) { g_attn_tc = on ? 1 : 0; }   // test / A-B hook

// extras of the stage-1 relative-position mode (relattn_tc.cu); all NULL / 0 for the GPT-2 attention
struct AttnTcRel {
  const void* bias;      // forward: [B*H][Tq][Tk] bf16 shifted position scores
  const void* biasT;     // backward: [B*H][Tk][Tq] bf16, transposed
  void* dbiasT;          // backward: receives dS^T in the same layout
  int rel;               // dropout = drop keys and renormalise; 1e-8 in the normaliser
};

int emo_attn_fwd_tc_launch_ex(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out, int64_t ld_o,
                              float* lse, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, const AttnTcRel* rel,
                              cudaStream_t s) {
  using namespace attn3;
  using namespace attn3::fwd;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = tcp::make_map_bt(&mq, q, (int64_t)H * HD, Tq, B, ld_q, BM))) return rc;
  if ((rc = tcp::make_map_bt(&mk, k, (int64_t)H * HD, Tk, B, ld_kv, BN))) return rc;
  if ((rc = tcp::make_map_bt(&mv, v, (int64_t)H * HD, Tk, B, ld_kv, BN))) return rc;
  Params p;
  p.out = (bf16*)out; p.ld_o = ld_o; p.lse = lse; p.Tq = Tq; p.Tk = Tk; p.H = H; p.nqt = (Tq + BM - 1) / BM;
  p.sl2 = scale * LOG2E; p.keep_scale = 1.f / (1.f - drop_p); p.drop_thr = emo_drop_thr(drop_p); p.seed = seed;
  p.bias = rel ? (const bf16*)rel->bias : nullptr;
  p.rel = rel ? rel->rel : 0;
  dim3 grid(p.nqt, B * H);
  const bool pair_hash = ((Tk & 1) == 0) && ((uint64_t)B * H * (uint64_t)Tq * (uint64_t)Tk + 512 < (1ull << 33));
  const int drop = p.drop_thr == 0 ? 0 : (pair_hash ? 1 : 2);
#define EMO_ATTN_FWD(R, D)                                                                                                   \
  do {                                                                                                                       \
    static bool configured = false;                                                                                          \
    if (!configured) {                                                                                                       \
      EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<R, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
      configured = true;                                                                                                     \
    }                                                                                                                        \
    attn_fwd_tc_kernel<R, D><<<grid, NT, SMEM_BYTES, s>>>(mq, mk, mv, p);                                                    \
  } while (0)
  if (p.rel) { if (drop == 0) EMO_ATTN_FWD(true, 0); else if (drop == 1) EMO_ATTN_FWD(true, 1); else EMO_ATTN_FWD(true, 2); }
  else { if (drop == 0) EMO_ATTN_FWD(false, 0); else if (drop == 1) EMO_ATTN_FWD(false, 1); else EMO_ATTN_FWD(false, 2); }
#undef EMO_ATTN_FWD
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
int emo_attn_fwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out, int64_t ld_o,
                           float* lse, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, cudaStream_t s) {
  return emo_attn_fwd_tc_launch_ex(q, k, v, ld_q, ld_kv, out, ld_o, lse, B, Tq, Tk, H, scale, drop_p, seed, nullptr, s);
}

// fp32 floats of workspace the backward core needs: dQ accumulator [B][Tq][H*64] + lse2 / dsum [B*H][Tq_pad]
int64_t emo_attn_bwd_tc_ws_floats(int B, int Tq, int H) {
  const int64_t Tq_pad = (int64_t)((Tq + attn3::BM - 1) / attn3::BM) * attn3::BM;
  return (int64_t)B * Tq * H * attn3::HD + 2 * (int64_t)B * H * Tq_pad;
}
int emo_attn_tc_configure_pool() {
  static bool done = false;
  if (done) return EMO_OK;
  int dev = 0;
  cudaMemPool_t pool;
  EMO_CHECK_CUDA(cudaGetDevice(&dev));
  EMO_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t keep = ~0ull;                        // stream-ordered workspaces are recycled, not returned to the OS
  EMO_CHECK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  done = true;
  return EMO_OK;
}
// everything of the backward except the final fp32 -> bf16 conversion of dQ: zeroes ws, row statistics, main kernel.
// On return (stream order) ws[0 .. B*Tq*H*64) holds dQ in fp32.
int emo_attn_bwd_tc_core(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* out, const void* dout,
                         int64_t ld_o, const float* lse, float* ws, void* dk, void* dv, int64_t ld_dkv, int B, int Tq, int Tk, int H,
                         float scale, float drop_p, uint64_t seed, const AttnTcRel* rel, cudaStream_t s) {
  using namespace attn3;
  using namespace attn3::bwd;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int nqt = (Tq + BM - 1) / BM, nkt = (Tk + BN - 1) / BN, Tq_pad = nqt * BM;
  const int64_t n_acc = (int64_t)B * Tq * H * HD, n_vec = (int64_t)B * H * Tq_pad;
  float* dqacc = ws;
  float* lse2 = ws + n_acc;
  float* dsum = lse2 + n_vec;
  EMO_CHECK_CUDA(cudaMemsetAsync(dqacc, 0, (size_t)n_acc * sizeof(float), s));
  attn_bwd_prep_kernel<<<(unsigned)((n_vec * 8 + 255) / 256), 256, 0, s>>>((const bf16*)out, (const bf16*)dout, ld_o, lse, lse2, dsum, Tq, Tq_pad, H, n_vec, scale);
  EMO_LAUNCH_CHECK();
  CUtensorMap mq, mk, mv, mdo, mdq;
  int rc;
  if ((rc = tcp::make_map_bt(&mq, q, (int64_t)H * HD, Tq, B, ld_q, BM))) return rc;
  if ((rc = tcp::make_map_bt(&mk, k, (int64_t)H * HD, Tk, B, ld_kv, BN))) return rc;
  if ((rc = tcp::make_map_bt(&mv, v, (int64_t)H * HD, Tk, B, ld_kv, BN))) return rc;
  if ((rc = tcp::make_map_bt(&mdo, dout, (int64_t)H * HD, Tq, B, ld_o, BM))) return rc;
  if ((rc = tcp::make_map_f32_bt(&mdq, dqacc, (int64_t)H * HD, Tq, B, 32, 16))) return rc;
  Params p;
  p.lse2 = lse2; p.dsum = dsum; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.ld_dkv = ld_dkv;
  p.Tq = Tq; p.Tk = Tk; p.Tq_pad = Tq_pad; p.H = H; p.nqt = nqt; p.scale = scale; p.sl2 = scale * LOG2E;
  p.keep_scale = 1.f / (1.f - drop_p); p.drop_thr = emo_drop_thr(drop_p); p.seed = seed;
  p.drop_fast = ((Tk & 1) == 0) && ((uint64_t)B * H * (uint64_t)Tq * (uint64_t)Tk < (1ull << 33));
  p.biasT = rel ? (const bf16*)rel->biasT : nullptr;
  p.dbiasT = rel ? (bf16*)rel->dbiasT : nullptr;
  p.rel = rel ? rel->rel : 0;
  { const char* e = getenv("EMO_ATTN_DBG_CLK"); p.dbg_clk = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
  dim3 grid(nkt, B * H);
  attn_bwd_tc_kernel<<<grid, NT, SMEM_BYTES, s>>>(mq, mk, mv, mdo, mdq, p);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
// dq (bf16, row stride ld_dq) = fp32 accumulator (+ optional dense bf16 addend)
int emo_attn_bwd_tc_convert(const float* dqacc, const void* add, void* dq, int64_t ld_dq, int B, int Tq, int H, cudaStream_t s) {
  using namespace attn3;
  const int64_t nconv = (int64_t)B * Tq * (H * HD / 8);
  bwd::attn_bwd_dq_convert_kernel<<<(unsigned)((nconv + 255) / 256), 256, 0, s>>>(dqacc, (const bf16*)add, (bf16*)dq, ld_dq, (int64_t)B * Tq, H * HD);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

int emo_attn_bwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* out,
                           const void* dout, int64_t ld_o, const float* lse, void* dq, void* dk, void* dv, int64_t ld_dq,
                           int64_t ld_dkv, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, cudaStream_t s) {
  int rc = emo_attn_tc_configure_pool();
  if (rc) return rc;
  float* ws = nullptr;
  EMO_CHECK_CUDA(cudaMallocAsync((void**)&ws, (size_t)emo_attn_bwd_tc_ws_floats(B, Tq, H) * sizeof(float), s));
  rc = emo_attn_bwd_tc_core(q, k, v, ld_q, ld_kv, out, dout, ld_o, lse, ws, dk, dv, ld_dkv, B, Tq, Tk, H, scale, drop_p, seed, nullptr, s);
  if (!rc) rc = emo_attn_bwd_tc_convert(ws, nullptr, dq, ld_dq, B, Tq, H, s);
  cudaError_t e2 = cudaFreeAsync(ws, s);
  if (rc) return rc;
  if (e2 != cudaSuccess) {
    emo_set_error("emo_attn_bwd (tcgen05): CUDA error %d (%s)", (int)e2, cudaGetErrorString(e2));
    return EMO_ERR_CUDA;
  }
  return EMO_OK;
}
