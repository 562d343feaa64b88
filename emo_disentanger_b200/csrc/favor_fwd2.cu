// FAVOR+ forward, register-resident formulation (bf16).  Same math and interface as favor_fwd_kernel
// (favor_kernels.cuh), different data flow: the block-GEMM version round-trips every intermediate (phi(q), the
// masked score tile, the bf16 state copy, the output) through shared memory between its small products and is bound
// by the shared-memory pipe (profiles/r01_favor_fwd_summary.txt: 37 % mma.sync pipe, stalls = short_scoreboard /
// mio_throttle).  Here a warp owns 16 query rows of the 64-token chunk end to end, flash-attention style:
//
//   U_q = X_q . Om        accumulators -> phi(q) re-packed IN REGISTERS as the A operand of the next products
//   S   = phi(q) phi(k)^T  accumulators -> causal mask -> re-packed in registers as the A operand of S . V'
//   O'  = S V' + phi(q) S'_prev ;  out = O'[:, :64] / (O'[:, 64] + eps)
//   S' += phi(k)^T V'      each warp owns 32 feature rows of the prefix state (fp32 accumulators, persistent)
//
// Only phi(k) (needed by every warp, as a B operand and transposed as an A operand) and the bf16 copy of the state go
// through shared memory; three block barriers per chunk instead of eleven.  4 warps per CTA, two CTAs per SM.
#include "common.cuh"

namespace favor2 {

constexpr int FE = 64, FM = 128, FV = 80, C = 64, NT_ = 128;   // head dim, features, padded value width, chunk, threads
constexpr int LD64 = 72, LD128 = 136, LD80 = 88;               // padded leading dims (bf16 elements): conflict-free ldmatrix
constexpr float F_EPS = 1e-6f, F_S2 = 0.125f, F_HALF_LOG_M = 2.4260151319598084f, K2 = 1.4426950408889634f;

struct Smem {
  bf16 xq[2][C][LD64];
  bf16 xk[2][C][LD64];
  bf16 xv[2][C][LD80];
  bf16 om[FE][LD64];
  bf16 pk[C][LD128];
  bf16 s[FM][LD80];
};
static_assert(2 * (sizeof(Smem) + 1024) <= 227 * 1024, "two CTAs per SM");

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cpa16(void* dst, const void* src, bool pred) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  int n = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

// [C x 64] rows of q / k / v (token stride ld) -> smem tiles; rows >= valid are zero-filled
template <int LD>
__device__ __forceinline__ void issue_tile(const bf16* src, int64_t ld, int valid, bf16 (*dst)[LD]) {
  for (int i = threadIdx.x; i < C * 8; i += NT_) {
    int row = i >> 3, part = i & 7;
    bool ok = row < valid;
    cpa16(&dst[row][part * 8], ok ? src + (int64_t)row * ld + part * 8 : src, ok);
  }
}

// U = X[16 rows of this warp] . Om  (16 x 64, fp32 accumulators) and the rows' exponent offsets
//   o = (|x|^2 s^2 / 2 + ln(128)/2) * log2(e)       for rows g and g + 8 of the warp's 16
__device__ __forceinline__ void project(const bf16 (*x)[LD64], const bf16 (*om)[LD64], int r0, float (&u)[8][4], float& o_lo, float& o_hi) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) u[j][i] = 0.f;
  float ss_lo = 0.f, ss_hi = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm4(a, &x[r0 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]);
    {  // a[0], a[2]: row g ; a[1], a[3]: row g + 8
      float f0, f1;
      unpack_bf16x2(a[0], f0, f1); ss_lo += f0 * f0 + f1 * f1;
      unpack_bf16x2(a[2], f0, f1); ss_lo += f0 * f0 + f1 * f1;
      unpack_bf16x2(a[1], f0, f1); ss_hi += f0 * f0 + f1 * f1;
      unpack_bf16x2(a[3], f0, f1); ss_hi += f0 * f0 + f1 * f1;
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {      // B = Om stored [K = e][N = f] (N contiguous) -> ldmatrix.trans
      uint32_t b[4];
      ldsm4t(b, &om[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][j * 8 + (lane >> 4) * 8]);
      mma16816(u[j], a, b[0], b[1]);
      mma16816(u[j + 1], a, b[2], b[3]);
    }
  }
  ss_lo += __shfl_xor_sync(0xffffffffu, ss_lo, 1); ss_lo += __shfl_xor_sync(0xffffffffu, ss_lo, 2);
  ss_hi += __shfl_xor_sync(0xffffffffu, ss_hi, 1); ss_hi += __shfl_xor_sync(0xffffffffu, ss_hi, 2);
  o_lo = (0.5f * F_S2 * ss_lo + F_HALF_LOG_M) * K2;
  o_hi = (0.5f * F_S2 * ss_hi + F_HALF_LOG_M) * K2;
}

__global__ void __launch_bounds__(NT_, 2)
favor_fwd2_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int64_t ld,
                  const float* __restrict__ omega, bf16* __restrict__ out, int64_t ld_out, float* __restrict__ den_out,
                  const float* __restrict__ state_in, float* __restrict__ state_out, float* __restrict__ seg_states,
                  int nseg, int seg_chunks, int Tlen, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const int bh = blockIdx.x / nseg, seg = blockIdx.x % nseg;
  const int b = bh / H, h = bh % H;
  const int64_t base = (int64_t)b * Tlen * ld + (int64_t)h * FE;
  const int64_t obase = (int64_t)b * Tlen * ld_out + (int64_t)h * FE;
  const int r0 = warp * 16;                 // this warp's query rows of a chunk
  const int f0 = warp * 32;                 // this warp's feature rows of the prefix state

  // Omega * 64^(-1/4) * log2(e) -> bf16 [e][f]
  {
    const float sc = 0.35355339059327373f * K2;
    for (int i = threadIdx.x; i < FE * FE / 4; i += NT_) {
      float4 w = __ldg(reinterpret_cast<const float4*>(omega) + i);
      bf16* d = &sm.om[(i * 4) / FE][(i * 4) % FE];
      *reinterpret_cast<uint32_t*>(d) = pack_bf16x2(w.x * sc, w.y * sc);
      *reinterpret_cast<uint32_t*>(d + 2) = pack_bf16x2(w.z * sc, w.w * sc);
    }
  }
  // ones column (and zero pad) of V' in both buffers; the copies only write columns 0..63
  for (int i = threadIdx.x; i < 2 * C * (FV - FE); i += NT_) {
    int bu = i / (C * (FV - FE)), r = (i / (FV - FE)) % C, c = FE + i % (FV - FE);
    sm.xv[bu][r][c] = __float2bfloat16_rn(c == FE ? 1.f : 0.f);
  }
  // prefix state of this warp's 32 feature rows: st[mt][j][i] <-> (row f0 + 16 mt + g (+8), col 8 j + 2 t4 (+1))
  float st[2][10][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) st[mt][j][i] = 0.f;
  const float* sin = nullptr;
  if (state_in) sin = state_in + (int64_t)bh * FM * FV;
  else if (seg_states && seg > 0) sin = seg_states + ((int64_t)bh * (nseg + 1) + seg) * FM * FV;
  if (sin) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const float2 lo = *reinterpret_cast<const float2*>(sin + (f0 + 16 * mt + g) * FV + 8 * j + 2 * t4);
        const float2 hi = *reinterpret_cast<const float2*>(sin + (f0 + 16 * mt + g + 8) * FV + 8 * j + 2 * t4);
        st[mt][j][0] = lo.x; st[mt][j][1] = lo.y; st[mt][j][2] = hi.x; st[mt][j][3] = hi.y;
      }
  }
  auto store_state_bf16 = [&]() {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        *reinterpret_cast<uint32_t*>(&sm.s[f0 + 16 * mt + g][8 * j + 2 * t4]) = pack_bf16x2(st[mt][j][0], st[mt][j][1]);
        *reinterpret_cast<uint32_t*>(&sm.s[f0 + 16 * mt + g + 8][8 * j + 2 * t4]) = pack_bf16x2(st[mt][j][2], st[mt][j][3]);
      }
  };
  store_state_bf16();

  const int t_begin = seg * seg_chunks * C;
  const int t_end = (t_begin + seg_chunks * C < Tlen) ? t_begin + seg_chunks * C : Tlen;
  if (t_begin < t_end) {
    const int valid0 = (Tlen - t_begin < C) ? (Tlen - t_begin) : C;
    issue_tile(q + base + (int64_t)t_begin * ld, ld, valid0, sm.xq[0]);
    issue_tile(k + base + (int64_t)t_begin * ld, ld, valid0, sm.xk[0]);
    issue_tile(v + base + (int64_t)t_begin * ld, ld, valid0, sm.xv[0]);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  int buf = 0;
  for (int t0 = t_begin; t0 < t_end; t0 += C, buf ^= 1) {
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // [S1] tiles of this chunk landed; the other buffers and s are free
    {
      const int tn = t0 + C;
      if (tn < t_end) {
        const int validn = (Tlen - tn < C) ? (Tlen - tn) : C;
        issue_tile(q + base + (int64_t)tn * ld, ld, validn, sm.xq[buf ^ 1]);
        issue_tile(k + base + (int64_t)tn * ld, ld, validn, sm.xk[buf ^ 1]);
        issue_tile(v + base + (int64_t)tn * ld, ld, validn, sm.xv[buf ^ 1]);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- phi(k) of this warp's 16 key rows -> shared (rows >= valid: zero) ----
    {
      float u[8][4], o_lo, o_hi;
      project(sm.xk[buf], sm.om, r0, u, o_lo, o_hi);
      const bool ok_lo = r0 + g < valid, ok_hi = r0 + g + 8 < valid;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = 8 * j + 2 * t4;
        *reinterpret_cast<uint32_t*>(&sm.pk[r0 + g][col]) = ok_lo ? pack_bf16x2(ex2(u[j][0] - o_lo), ex2(u[j][1] - o_lo)) : 0u;
        *reinterpret_cast<uint32_t*>(&sm.pk[r0 + g][FE + col]) = ok_lo ? pack_bf16x2(ex2(-u[j][0] - o_lo), ex2(-u[j][1] - o_lo)) : 0u;
        *reinterpret_cast<uint32_t*>(&sm.pk[r0 + g + 8][col]) = ok_hi ? pack_bf16x2(ex2(u[j][2] - o_hi), ex2(u[j][3] - o_hi)) : 0u;
        *reinterpret_cast<uint32_t*>(&sm.pk[r0 + g + 8][FE + col]) = ok_hi ? pack_bf16x2(ex2(-u[j][2] - o_hi), ex2(-u[j][3] - o_hi)) : 0u;
      }
    }
    // ---- phi(q) of this warp's 16 query rows, kept as A fragments: qa[ks] covers features 16 ks .. 16 ks + 15 ----
    uint32_t qa[8][4];
    {
      float u[8][4], o_lo, o_hi;
      project(sm.xq[buf], sm.om, r0, u, o_lo, o_hi);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                 // accumulator tiles (2 ks, 2 ks + 1) -> one k16 A fragment
        qa[ks][0] = pack_bf16x2(ex2(u[2 * ks][0] - o_lo), ex2(u[2 * ks][1] - o_lo));
        qa[ks][1] = pack_bf16x2(ex2(u[2 * ks][2] - o_hi), ex2(u[2 * ks][3] - o_hi));
        qa[ks][2] = pack_bf16x2(ex2(u[2 * ks + 1][0] - o_lo), ex2(u[2 * ks + 1][1] - o_lo));
        qa[ks][3] = pack_bf16x2(ex2(u[2 * ks + 1][2] - o_hi), ex2(u[2 * ks + 1][3] - o_hi));
        qa[4 + ks][0] = pack_bf16x2(ex2(-u[2 * ks][0] - o_lo), ex2(-u[2 * ks][1] - o_lo));
        qa[4 + ks][1] = pack_bf16x2(ex2(-u[2 * ks][2] - o_hi), ex2(-u[2 * ks][3] - o_hi));
        qa[4 + ks][2] = pack_bf16x2(ex2(-u[2 * ks + 1][0] - o_lo), ex2(-u[2 * ks + 1][1] - o_lo));
        qa[4 + ks][3] = pack_bf16x2(ex2(-u[2 * ks + 1][2] - o_hi), ex2(-u[2 * ks + 1][3] - o_hi));
      }
    }
    __syncthreads();                                   // [S2] phi(k) tile complete; everybody is done with xq[buf]
    // ---- O' = tril(phi(q) phi(k)^T) V' + phi(q) S' ----
    float o[10][4];
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                   // 16 keys at a time; keys beyond this warp's rows are masked anyway
      if (kk <= warp) {
        float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {               // B = phi(k) stored [N = token][K = feature] (K contiguous)
          uint32_t bfr[4];
          ldsm4(bfr, &sm.pk[kk * 16 + (lane & 7) + (lane >> 4) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
          mma16816(sc[0], qa[ks], bfr[0], bfr[1]);
          mma16816(sc[1], qa[ks], bfr[2], bfr[3]);
        }
        uint32_t pa[4];
        if (kk == warp) {                              // diagonal block: keep key <= query
          const int c0 = 2 * t4, c1 = 8 + 2 * t4;      // key columns of the two tiles (relative), rows g / g + 8
          sc[0][0] = (c0 <= g) ? sc[0][0] : 0.f;         sc[0][1] = (c0 + 1 <= g) ? sc[0][1] : 0.f;
          sc[0][2] = (c0 <= g + 8) ? sc[0][2] : 0.f;     sc[0][3] = (c0 + 1 <= g + 8) ? sc[0][3] : 0.f;
          sc[1][0] = (c1 <= g) ? sc[1][0] : 0.f;         sc[1][1] = (c1 + 1 <= g) ? sc[1][1] : 0.f;
          sc[1][2] = (c1 <= g + 8) ? sc[1][2] : 0.f;     sc[1][3] = (c1 + 1 <= g + 8) ? sc[1][3] : 0.f;
        }
        pa[0] = pack_bf16x2(sc[0][0], sc[0][1]);
        pa[1] = pack_bf16x2(sc[0][2], sc[0][3]);
        pa[2] = pack_bf16x2(sc[1][0], sc[1][1]);
        pa[3] = pack_bf16x2(sc[1][2], sc[1][3]);
#pragma unroll
        for (int j = 0; j < 10; j += 2) {              // B = V' stored [K = token][N = 80] (N contiguous) -> .trans
          uint32_t bfr[4];
          ldsm4t(bfr, &sm.xv[buf][kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][j * 8 + (lane >> 4) * 8]);
          mma16816(o[j], pa, bfr[0], bfr[1]);
          mma16816(o[j + 1], pa, bfr[2], bfr[3]);
        }
      }
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {                   // + phi(q) S'_prev,  S' bf16 stored [K = feature][N = 80]
#pragma unroll
      for (int j = 0; j < 10; j += 2) {
        uint32_t bfr[4];
        ldsm4t(bfr, &sm.s[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][j * 8 + (lane >> 4) * 8]);
        mma16816(o[j], qa[ks], bfr[0], bfr[1]);
        mma16816(o[j + 1], qa[ks], bfr[2], bfr[3]);
      }
    }
    // ---- normalise by the ones column (col 64 = tile 8, elements 0 / 2 of the t4 == 0 lanes) and stage the rows ----
    {
      const float d_lo = __shfl_sync(0xffffffffu, o[8][0], lane & ~3) + F_EPS;
      const float d_hi = __shfl_sync(0xffffffffu, o[8][2], lane & ~3) + F_EPS;
      const float i_lo = 1.f / d_lo, i_hi = 1.f / d_hi;
      bf16 (*stage)[LD64] = sm.xq[buf];                // the q tile is dead (S2): output staging
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(&stage[r0 + g][8 * j + 2 * t4]) = pack_bf16x2(o[j][0] * i_lo, o[j][1] * i_lo);
        *reinterpret_cast<uint32_t*>(&stage[r0 + g + 8][8 * j + 2 * t4]) = pack_bf16x2(o[j][2] * i_hi, o[j][3] * i_hi);
      }
      if (den_out && t4 == 0) {
        if (r0 + g < valid) den_out[((int64_t)b * Tlen + t0 + r0 + g) * H + h] = d_lo;
        if (r0 + g + 8 < valid) den_out[((int64_t)b * Tlen + t0 + r0 + g + 8) * H + h] = d_hi;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {                    // this warp's 16 rows x 128 bytes, 16-byte stores
        const int idx = lane + 32 * i, row = r0 + (idx >> 3), part = idx & 7;
        if (row < valid)
          *reinterpret_cast<uint4*>(out + obase + (int64_t)(t0 + row) * ld_out + part * 8) = *reinterpret_cast<const uint4*>(&stage[row][part * 8]);
      }
    }
    // ---- S' += phi(k)^T V' for this warp's 32 feature rows ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                   // 16 tokens per step
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {                 // A = phi(k)^T: stored [K = token][M = feature] (M contiguous) -> .trans
        const int mat = lane >> 3;
        ldsm4t(a[mt], &sm.pk[kk * 16 + (lane & 7) + (mat >> 1) * 8][f0 + 16 * mt + (mat & 1) * 8]);
      }
#pragma unroll
      for (int j = 0; j < 10; j += 2) {
        uint32_t bfr[4];
        ldsm4t(bfr, &sm.xv[buf][kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][j * 8 + (lane >> 4) * 8]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(st[mt][j], a[mt], bfr[0], bfr[1]);
          mma16816(st[mt][j + 1], a[mt], bfr[2], bfr[3]);
        }
      }
    }
    __syncthreads();                                   // [S3] every warp has read s (and pk, xv[buf]) of this chunk
    store_state_bf16();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  auto store_state_f32 = [&](float* dst) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        *reinterpret_cast<float2*>(dst + (f0 + 16 * mt + g) * FV + 8 * j + 2 * t4) = make_float2(st[mt][j][0], st[mt][j][1]);
        *reinterpret_cast<float2*>(dst + (f0 + 16 * mt + g + 8) * FV + 8 * j + 2 * t4) = make_float2(st[mt][j][2], st[mt][j][3]);
      }
  };
  if (seg == nseg - 1) {
    if (state_out) store_state_f32(state_out + (int64_t)bh * FM * FV);
    if (seg_states) store_state_f32(seg_states + ((int64_t)bh * (nseg + 1) + nseg) * FM * FV);
  }
  if (seg_states && seg == 0) {
    float* so = seg_states + (int64_t)bh * (nseg + 1) * FM * FV;
    for (int i = threadIdx.x; i < FM * FV; i += NT_) so[i] = 0.f;
  }
}

}  // namespace favor2

int emo_favor_fwd2_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, void* out,
                          int64_t ld_out, float* den, const float* state_in, float* state_out, float* seg_states,
                          int nseg, int sc, int B, int T_, int H, cudaStream_t s) {
  using namespace favor2;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    configured = true;
  }
  favor_fwd2_kernel<<<B * H * nseg, NT_, sizeof(Smem), s>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, ld, omega, (bf16*)out,
                                                           ld_out, den, state_in, state_out, seg_states, nseg, sc, T_, H);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
