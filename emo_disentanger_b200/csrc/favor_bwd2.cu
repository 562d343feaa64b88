// FAVOR+ backward, register-resident formulation (bf16).  Same math, interface and workspace layout as
// favor_bwd_kernel (favor_kernels.cuh); different data flow.  The block-GEMM version writes every intermediate (the two
// masked [64 x 64] tiles, d phi * phi, du, the staged outputs) to shared memory between its eleven small products and
// is bound by the shared-memory pipe.  Here the 8 warps of a CTA split into a query side and a key side, and each warp
// carries 16 token rows of the 64-token chunk through its whole chain in registers:
//
//   query warps (0-3), rows i:   P   = tril(G V'^T)                   accumulators -> A fragments
//                                dPq = P phi(k) + G S_prev^T          -> * phi(q) -> du, do in registers
//                                dq  = du Om^T / k2 + do s^2 x_q      -> staged in place of x_q, 16-byte stores
//                                R  += phi(q)^T G                     (each warp owns 32 feature rows, fp32, persistent)
//                                phi(k) R                             (half of dv, handed to the key side in fp32)
//   key warps (4-7), rows j:     S  -= phi(k)^T V'                    (rolled back to the chunk start, 32 rows per warp, kept as -S)
//                                P^T = triu(V' G^T);  dPk = P^T phi(q) + V' R^T  -> dk like dq
//                                A^T = triu(phi(k) phi(q)^T);  dv = A^T G + [phi(k) R]
//
// Shared memory carries only what another warp reads: phi(q), phi(k), G, the bf16 copies of S and R, and the dv
// partial.  Two block barriers per chunk; the hand-overs between the two groups (s ready, r free, partial ready) are
// arrive / wait pairs so neither group stalls on work it does not depend on.
#include "common.cuh"

namespace favor2b {

constexpr int FE = 64, FM = 128, FV = 80, C = 64, NT_ = 256;
constexpr int LD64 = 72, LD128 = 136, LD80 = 88;
constexpr float F_S2 = 0.125f, F_HALF_LOG_M = 2.4260151319598084f, K2 = 1.4426950408889634f, KINV = 0.6931471805599453f;

struct Smem {
  bf16 xq[2][C][LD64];
  bf16 xk[2][C][LD64];
  bf16 xv[2][C][LD80];
  bf16 om[FE][LD64];
  bf16 pq[C][LD128];
  bf16 pk[C][LD128];
  bf16 g[C][LD80];
  bf16 s[FM][LD80];
  bf16 r[FM][LD80];
  float4 dvp[4][8][32];            // phi(k) R partial of dv: [key warp][n-tile][lane] in accumulator layout
  bf16 ro[2][C][LD64];             // out / dout rows and den of the chunk (prefetched one chunk ahead like q, k, v)
  bf16 rd[2][C][LD64];
  float den[2][C];
};
static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "one CTA per SM");

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
__device__ __forceinline__ void block_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// producer / consumer barriers between the two warp groups (128 arriving + 128 waiting threads)
template <int ID> __device__ __forceinline__ void bar_arrive() { asm volatile("bar.arrive %0, 256;" ::"n"(ID) : "memory"); }
template <int ID> __device__ __forceinline__ void bar_wait() { asm volatile("bar.sync %0, 256;" ::"n"(ID) : "memory"); }
constexpr int BAR_S = 2, BAR_R = 3, BAR_DV = 4;   // s written | r no longer read | dv partial written

template <int LD>
__device__ __forceinline__ void issue_tile(const bf16* src, int64_t ld, int valid, bf16 (*dst)[LD]) {
  for (int i = threadIdx.x; i < C * 8; i += NT_) {
    int row = i >> 3, part = i & 7;
    bool ok = row < valid;
    cpa16(&dst[row][part * 8], ok ? src + (int64_t)row * ld + part * 8 : src, ok);
  }
}

// A-operand fragment addresses of 16 rows starting at r0 (row-major, K contiguous), k-step ks
#define A_ADDR(t, r0, ks) &(t)[(r0) + (lane & 7) + ((lane >> 3) & 1) * 8][(ks) * 16 + (lane >> 4) * 8]
// B operand stored [N][K] (K contiguous): two n-tiles starting at n0, k-step ks      (plain ldmatrix)
#define BNK_ADDR(t, n0, ks) &(t)[(n0) + (lane & 7) + (lane >> 4) * 8][(ks) * 16 + ((lane >> 3) & 1) * 8]
// B operand stored [K][N] (N contiguous): k rows starting at k0, two n-tiles at n0   (ldmatrix.trans)
#define BKN_ADDR(t, k0, n0) &(t)[(k0) + (lane & 7) + ((lane >> 3) & 1) * 8][(n0) + (lane >> 4) * 8]

// phi of this warp's 16 rows of x: packed bf16 pairs.  ph[ks] is at once the k16 A fragment over features
// 16 ks .. 16 ks + 15 and the accumulator-layout tiles (2 ks, 2 ks + 1): [0] tile 2ks row g, [1] tile 2ks row g + 8,
// [2] tile 2ks+1 row g, [3] tile 2ks+1 row g + 8.  ks 0..3: exp(+u - o), ks 4..7: exp(-u - o).  Also written to `dst`.
__device__ __forceinline__ void phi16(const bf16 (*x)[LD64], const bf16 (*om)[LD64], int r0, int valid, uint32_t (&ph)[8][4],
                                      bf16 (*dst)[LD128]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  float u[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) u[j][i] = 0.f;
  float ss_lo = 0.f, ss_hi = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm4(a, A_ADDR(x, r0, ks));
    {
      float f0, f1;
      unpack_bf16x2(a[0], f0, f1); ss_lo += f0 * f0 + f1 * f1;
      unpack_bf16x2(a[2], f0, f1); ss_lo += f0 * f0 + f1 * f1;
      unpack_bf16x2(a[1], f0, f1); ss_hi += f0 * f0 + f1 * f1;
      unpack_bf16x2(a[3], f0, f1); ss_hi += f0 * f0 + f1 * f1;
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {      // B = Om stored [K = e][N = f]
      uint32_t b[4];
      ldsm4t(b, BKN_ADDR(om, ks * 16, j * 8));
      mma16816(u[j], a, b[0], b[1]);
      mma16816(u[j + 1], a, b[2], b[3]);
    }
  }
  ss_lo += __shfl_xor_sync(0xffffffffu, ss_lo, 1); ss_lo += __shfl_xor_sync(0xffffffffu, ss_lo, 2);
  ss_hi += __shfl_xor_sync(0xffffffffu, ss_hi, 1); ss_hi += __shfl_xor_sync(0xffffffffu, ss_hi, 2);
  const float o_lo = (0.5f * F_S2 * ss_lo + F_HALF_LOG_M) * K2;
  const float o_hi = (0.5f * F_S2 * ss_hi + F_HALF_LOG_M) * K2;
  const bool ok_lo = r0 + g < valid, ok_hi = r0 + g + 8 < valid;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * ks + h;
      ph[ks][2 * h] = ok_lo ? pack_bf16x2(ex2(u[j][0] - o_lo), ex2(u[j][1] - o_lo)) : 0u;
      ph[ks][2 * h + 1] = ok_hi ? pack_bf16x2(ex2(u[j][2] - o_hi), ex2(u[j][3] - o_hi)) : 0u;
      ph[4 + ks][2 * h] = ok_lo ? pack_bf16x2(ex2(-u[j][0] - o_lo), ex2(-u[j][1] - o_lo)) : 0u;
      ph[4 + ks][2 * h + 1] = ok_hi ? pack_bf16x2(ex2(-u[j][2] - o_hi), ex2(-u[j][3] - o_hi)) : 0u;
    }
  }
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = 16 * ks + 8 * h + 2 * t4;
      *reinterpret_cast<uint32_t*>(&dst[r0 + g][col]) = ph[ks][2 * h];
      *reinterpret_cast<uint32_t*>(&dst[r0 + g + 8][col]) = ph[ks][2 * h + 1];
    }
}

// d (16 x 128 accumulators of d phi) -> dx rows: w = d * phi ; du = w+ - w- ; do = -sum w ;
// dx = du Om^T / k2 + do s^2 x, written over this warp's rows of x (same lane reads and writes each element) and
// copied out with 16-byte stores.
__device__ __forceinline__ void finish_dx(float (&d)[16][4], const uint32_t (&ph)[8][4], bf16 (*x)[LD64], const bf16 (*om)[LD64],
                                          int r0, int valid, bf16* __restrict__ dst, int64_t ld_d) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float p0, p1, p2, p3;
    unpack_bf16x2(ph[j >> 1][(j & 1) * 2], p0, p1);
    unpack_bf16x2(ph[j >> 1][(j & 1) * 2 + 1], p2, p3);
    d[j][0] *= p0; d[j][1] *= p1; d[j][2] *= p2; d[j][3] *= p3;
    s_lo += d[j][0] + d[j][1];
    s_hi += d[j][2] + d[j][3];
  }
  s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
  s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
  uint32_t da[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * ks + h;
      da[ks][2 * h] = pack_bf16x2(d[j][0] - d[j + 8][0], d[j][1] - d[j + 8][1]);
      da[ks][2 * h + 1] = pack_bf16x2(d[j][2] - d[j + 8][2], d[j][3] - d[j + 8][3]);
    }
  }
  float xx[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) xx[j][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)          // K = feature f, N = e ; B = Om^T = om[e][f]: [N][K]
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      uint32_t b[4];
      ldsm4(b, BNK_ADDR(om, j * 8, ks));
      mma16816(xx[j], da[ks], b[0], b[1]);
      mma16816(xx[j + 1], da[ks], b[2], b[3]);
    }
  const float c_lo = -s_lo * F_S2, c_hi = -s_hi * F_S2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t* p_lo = reinterpret_cast<uint32_t*>(&x[r0 + g][8 * j + 2 * t4]);
    uint32_t* p_hi = reinterpret_cast<uint32_t*>(&x[r0 + g + 8][8 * j + 2 * t4]);
    float x0, x1, x2, x3;
    unpack_bf16x2(*p_lo, x0, x1);
    unpack_bf16x2(*p_hi, x2, x3);
    *p_lo = pack_bf16x2(xx[j][0] * KINV + c_lo * x0, xx[j][1] * KINV + c_lo * x1);
    *p_hi = pack_bf16x2(xx[j][2] * KINV + c_hi * x2, xx[j][3] * KINV + c_hi * x3);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + 32 * i, row = r0 + (idx >> 3), part = idx & 7;
    if (row < valid) *reinterpret_cast<uint4*>(dst + (int64_t)row * ld_d + part * 8) = *reinterpret_cast<const uint4*>(&x[row][part * 8]);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(NT_, 1)
favor_bwd2_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int64_t ld,
                  const float* __restrict__ omega, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                  int64_t ld_out, const float* __restrict__ den_in, const float* __restrict__ seg_states,
                  const float* __restrict__ seg_rstates, int nseg, int seg_chunks, int fwd_nseg, int ratio,
                  bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv, int64_t ld_d, int Tlen, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const bool qside = warp < 4;
  const int w4 = warp & 3;
  const int r0 = w4 * 16;                   // this warp's token rows of a chunk
  const int f0 = w4 * 32;                   // this warp's feature rows of its state (R on the query side, S on the key side)
  const int bh = blockIdx.x / nseg, seg = blockIdx.x % nseg;
  const int b = bh / H, h = bh % H;
  const int64_t base = (int64_t)b * Tlen * ld + (int64_t)h * FE;
  const int64_t obase = (int64_t)b * Tlen * ld_out + (int64_t)h * FE;
  const int64_t dbase = (int64_t)b * Tlen * ld_d + (int64_t)h * FE;
  const int nchunk = (Tlen + C - 1) / C;
  const int c_begin = seg * seg_chunks;
  const int c_end = (c_begin + seg_chunks < nchunk) ? c_begin + seg_chunks : nchunk;
  if (c_begin >= c_end) return;

  auto issue_qkv = [&](int c, int buf) {
    if (c >= c_begin) {
      const int t0 = c * C;
      const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
      issue_tile(q + base + (int64_t)t0 * ld, ld, valid, sm.xq[buf]);
      issue_tile(k + base + (int64_t)t0 * ld, ld, valid, sm.xk[buf]);
      issue_tile(v + base + (int64_t)t0 * ld, ld, valid, sm.xv[buf]);
      issue_tile(out + obase + (int64_t)t0 * ld_out, ld_out, valid, sm.ro[buf]);
      issue_tile(dout + obase + (int64_t)t0 * ld_out, ld_out, valid, sm.rd[buf]);
      if (threadIdx.x < C) {
        const bool ok = threadIdx.x < valid;
        const float* src = den_in + ((int64_t)b * Tlen + t0 + (ok ? threadIdx.x : 0)) * H + h;
        uint32_t d = (uint32_t)__cvta_generic_to_shared(&sm.den[buf][threadIdx.x]);
        int n = ok ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_qkv(c_end - 1, 0);
  {
    const float sc = 0.35355339059327373f * K2;
    for (int i = threadIdx.x; i < FE * FE / 4; i += NT_) {
      float4 w = __ldg(reinterpret_cast<const float4*>(omega) + i);
      bf16* d = &sm.om[(i * 4) / FE][(i * 4) % FE];
      *reinterpret_cast<uint32_t*>(d) = pack_bf16x2(w.x * sc, w.y * sc);
      *reinterpret_cast<uint32_t*>(d + 2) = pack_bf16x2(w.z * sc, w.w * sc);
    }
  }
  for (int i = threadIdx.x; i < 2 * C * (FV - FE); i += NT_) {
    int bu = i / (C * (FV - FE)), r = (i / (FV - FE)) % C, c = FE + i % (FV - FE);
    sm.xv[bu][r][c] = __float2bfloat16_rn(c == FE ? 1.f : 0.f);
  }
  // this warp's 32 rows of its state: query warps carry R (reverse state), key warps carry S (prefix state)
  float st[2][10][4];
  {
    const float* src = nullptr;
    if (qside) {
      if (nseg > 1) src = seg_rstates + ((int64_t)bh * (nseg + 1) + seg) * FM * FV;
    } else {
      int slot = (seg + 1) * ratio;
      if (slot > fwd_nseg) slot = fwd_nseg;
      src = seg_states + ((int64_t)bh * (fwd_nseg + 1) + slot) * FM * FV;
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        float2 lo = make_float2(0.f, 0.f), hi = lo;
        if (src) {
          lo = *reinterpret_cast<const float2*>(src + (f0 + 16 * mt + g) * FV + 8 * j + 2 * t4);
          hi = *reinterpret_cast<const float2*>(src + (f0 + 16 * mt + g + 8) * FV + 8 * j + 2 * t4);
        }
        // the key side carries -S, so that the roll-back S -= phi(k)^T V' is a plain accumulation
        const float sg = qside ? 1.f : -1.f;
        st[mt][j][0] = sg * lo.x; st[mt][j][1] = sg * lo.y; st[mt][j][2] = sg * hi.x; st[mt][j][3] = sg * hi.y;
      }
  }
  auto store_state_bf16 = [&](bf16 (*dst)[LD80], float sg) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        *reinterpret_cast<uint32_t*>(&dst[f0 + 16 * mt + g][8 * j + 2 * t4]) = pack_bf16x2(sg * st[mt][j][0], sg * st[mt][j][1]);
        *reinterpret_cast<uint32_t*>(&dst[f0 + 16 * mt + g + 8][8 * j + 2 * t4]) = pack_bf16x2(sg * st[mt][j][2], sg * st[mt][j][3]);
      }
  };
  if (qside) store_state_bf16(sm.r, 1.f);

  int buf = 0;
  for (int c = c_end - 1; c >= c_begin; --c, buf ^= 1) {
    const int t0 = c * C;
    const int valid = (Tlen - t0 < C) ? (Tlen - t0) : C;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    block_bar();                                       // [S1] tiles landed; r holds R of the later chunks; other buffers free
    issue_qkv(c - 1, buf ^ 1);
    bf16 (*xv)[LD80] = sm.xv[buf];

    if (qside) {
      // ================================ query side ================================
      uint32_t ph[8][4];
      phi16(sm.xq[buf], sm.om, r0, valid, ph, sm.pq);
      {  // G rows of this warp: lane -> (row r0 + lane / 2, 32-column half lane % 2); rows >= valid are zero tiles
        const int row = r0 + (lane >> 1), half = lane & 1;
        const float inv = row < valid ? 1.f / sm.den[buf][row] : 0.f;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 od = *reinterpret_cast<const uint4*>(&sm.ro[buf][row][half * 32 + i * 8]);
          const uint4 dd = *reinterpret_cast<const uint4*>(&sm.rd[buf][row][half * 32 + i * 8]);
          const uint32_t* po = reinterpret_cast<const uint32_t*>(&od);
          const uint32_t* pd = reinterpret_cast<const uint32_t*>(&dd);
          uint4 gq;
          uint32_t* pg = reinterpret_cast<uint32_t*>(&gq);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float o0, o1, d0, d1;
            unpack_bf16x2(po[e], o0, o1);
            unpack_bf16x2(pd[e], d0, d1);
            dot += o0 * d0 + o1 * d1;
            pg[e] = pack_bf16x2(d0 * inv, d1 * inv);
          }
          *reinterpret_cast<uint4*>(&sm.g[row][half * 32 + i * 8]) = gq;
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        uint4 tail = make_uint4(0, 0, 0, 0);
        if (half == 0) tail.x = pack_bf16x2(-dot * inv, 0.f);
        *reinterpret_cast<uint4*>(&sm.g[row][FE + half * 8]) = tail;
      }
      block_bar();                                     // [S2] phi(q), phi(k), G complete
      uint32_t ga[5][4];                               // G rows of this warp as A fragments
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) ldsm4(ga[ks], A_ADDR(sm.g, r0, ks));
      float d[16][4];
#pragma unroll
      for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk <= w4) {
          float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {             // B = V' stored [N = key][K = 80]
            uint32_t bfr[4];
            ldsm4(bfr, BNK_ADDR(xv, kk * 16, ks));
            mma16816(sc[0], ga[ks], bfr[0], bfr[1]);
            mma16816(sc[1], ga[ks], bfr[2], bfr[3]);
          }
          if (kk == w4) {                              // keep key <= query
            const int c0 = 2 * t4, c1 = 8 + 2 * t4;
            sc[0][0] = (c0 <= g) ? sc[0][0] : 0.f;         sc[0][1] = (c0 + 1 <= g) ? sc[0][1] : 0.f;
            sc[0][2] = (c0 <= g + 8) ? sc[0][2] : 0.f;     sc[0][3] = (c0 + 1 <= g + 8) ? sc[0][3] : 0.f;
            sc[1][0] = (c1 <= g) ? sc[1][0] : 0.f;         sc[1][1] = (c1 + 1 <= g) ? sc[1][1] : 0.f;
            sc[1][2] = (c1 <= g + 8) ? sc[1][2] : 0.f;     sc[1][3] = (c1 + 1 <= g + 8) ? sc[1][3] : 0.f;
          }
          uint32_t pa[4];
          pa[0] = pack_bf16x2(sc[0][0], sc[0][1]);
          pa[1] = pack_bf16x2(sc[0][2], sc[0][3]);
          pa[2] = pack_bf16x2(sc[1][0], sc[1][1]);
          pa[3] = pack_bf16x2(sc[1][2], sc[1][3]);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {            // B = phi(k) stored [K = key][N = feature]
            uint32_t bfr[4];
            ldsm4t(bfr, BKN_ADDR(sm.pk, kk * 16, j * 8));
            mma16816(d[j], pa, bfr[0], bfr[1]);
            mma16816(d[j + 1], pa, bfr[2], bfr[3]);
          }
        }
      }
      bar_wait<BAR_S>();                               // s holds the prefix state at the start of this chunk
#pragma unroll
      for (int ks = 0; ks < 5; ++ks)                   // + G S^T : B = S stored [N = feature][K = 80]
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          uint32_t bfr[4];
          ldsm4(bfr, BNK_ADDR(sm.s, j * 8, ks));
          mma16816(d[j], ga[ks], bfr[0], bfr[1]);
          mma16816(d[j + 1], ga[ks], bfr[2], bfr[3]);
        }
      finish_dx(d, ph, sm.xq[buf], sm.om, r0, valid, dq + dbase + (int64_t)t0 * ld_d, ld_d);
      {  // the phi(k) R half of dv for key rows r0.. (balances the two warp groups); handed over in fp32
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint32_t a[4];
          ldsm4(a, A_ADDR(sm.pk, r0, ks));
#pragma unroll
          for (int j = 0; j < 8; j += 2) {             // B = R stored [K = feature][N = 80]
            uint32_t bfr[4];
            ldsm4t(bfr, BKN_ADDR(sm.r, ks * 16, j * 8));
            mma16816(o[j], a, bfr[0], bfr[1]);
            mma16816(o[j + 1], a, bfr[2], bfr[3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.dvp[w4][j][lane] = make_float4(o[j][0], o[j][1], o[j][2], o[j][3]);
      }
      bar_arrive<BAR_DV>();
      // R += phi(q)^T G for this warp's 32 feature rows
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int mat = lane >> 3;
          ldsm4t(a[mt], &sm.pq[kk * 16 + (lane & 7) + (mat >> 1) * 8][f0 + 16 * mt + (mat & 1) * 8]);
        }
#pragma unroll
        for (int j = 0; j < 10; j += 2) {
          uint32_t bfr[4];
          ldsm4t(bfr, BKN_ADDR(sm.g, kk * 16, j * 8));
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(st[mt][j], a[mt], bfr[0], bfr[1]);
            mma16816(st[mt][j + 1], a[mt], bfr[2], bfr[3]);
          }
        }
      }
      bar_wait<BAR_R>();                               // the key side is done with r
      store_state_bf16(sm.r, 1.f);
    } else {
      // ================================= key side =================================
      uint32_t ph[8][4];
      phi16(sm.xk[buf], sm.om, r0, valid, ph, sm.pk);
      block_bar();                                     // [S2]
      // (-S) += phi(k)^T V' (roll back to the start of this chunk)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int mat = lane >> 3;
          ldsm4t(a[mt], &sm.pk[kk * 16 + (lane & 7) + (mat >> 1) * 8][f0 + 16 * mt + (mat & 1) * 8]);
        }
#pragma unroll
        for (int j = 0; j < 10; j += 2) {
          uint32_t bfr[4];
          ldsm4t(bfr, BKN_ADDR(xv, kk * 16, j * 8));
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(st[mt][j], a[mt], bfr[0], bfr[1]);
            mma16816(st[mt][j + 1], a[mt], bfr[2], bfr[3]);
          }
        }
      }
      store_state_bf16(sm.s, -1.f);
      bar_arrive<BAR_S>();
      uint32_t va[5][4];                               // V' rows of this warp as A fragments
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) ldsm4(va[ks], A_ADDR(xv, r0, ks));
      {
        float d[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk >= w4) {
            float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {           // P^T block = V' G^T : B = G stored [N = query][K = 80]
              uint32_t bfr[4];
              ldsm4(bfr, BNK_ADDR(sm.g, kk * 16, ks));
              mma16816(sc[0], va[ks], bfr[0], bfr[1]);
              mma16816(sc[1], va[ks], bfr[2], bfr[3]);
            }
            if (kk == w4) {                            // keep query >= key (rows = keys g / g + 8, columns = queries)
              const int c0 = 2 * t4, c1 = 8 + 2 * t4;
              sc[0][0] = (c0 >= g) ? sc[0][0] : 0.f;         sc[0][1] = (c0 + 1 >= g) ? sc[0][1] : 0.f;
              sc[0][2] = (c0 >= g + 8) ? sc[0][2] : 0.f;     sc[0][3] = (c0 + 1 >= g + 8) ? sc[0][3] : 0.f;
              sc[1][0] = (c1 >= g) ? sc[1][0] : 0.f;         sc[1][1] = (c1 + 1 >= g) ? sc[1][1] : 0.f;
              sc[1][2] = (c1 >= g + 8) ? sc[1][2] : 0.f;     sc[1][3] = (c1 + 1 >= g + 8) ? sc[1][3] : 0.f;
            }
            uint32_t pa[4];
            pa[0] = pack_bf16x2(sc[0][0], sc[0][1]);
            pa[1] = pack_bf16x2(sc[0][2], sc[0][3]);
            pa[2] = pack_bf16x2(sc[1][0], sc[1][1]);
            pa[3] = pack_bf16x2(sc[1][2], sc[1][3]);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {          // B = phi(q) stored [K = query][N = feature]
              uint32_t bfr[4];
              ldsm4t(bfr, BKN_ADDR(sm.pq, kk * 16, j * 8));
              mma16816(d[j], pa, bfr[0], bfr[1]);
              mma16816(d[j + 1], pa, bfr[2], bfr[3]);
            }
          }
        }
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)                 // + V' R^T : B = R stored [N = feature][K = 80]
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            uint32_t bfr[4];
            ldsm4(bfr, BNK_ADDR(sm.r, j * 8, ks));
            mma16816(d[j], va[ks], bfr[0], bfr[1]);
            mma16816(d[j + 1], va[ks], bfr[2], bfr[3]);
          }
        bar_arrive<BAR_R>();
        finish_dx(d, ph, sm.xk[buf], sm.om, r0, valid, dk + dbase + (int64_t)t0 * ld_d, ld_d);
      }
      // dv = triu(phi(k) phi(q)^T) G + phi(k) R  (64 value columns)
      {
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk >= w4) {
            float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {           // B = phi(q) stored [N = query][K = feature]
              uint32_t bfr[4];
              ldsm4(bfr, BNK_ADDR(sm.pq, kk * 16, ks));
              mma16816(sc[0], ph[ks], bfr[0], bfr[1]);
              mma16816(sc[1], ph[ks], bfr[2], bfr[3]);
            }
            if (kk == w4) {
              const int c0 = 2 * t4, c1 = 8 + 2 * t4;
              sc[0][0] = (c0 >= g) ? sc[0][0] : 0.f;         sc[0][1] = (c0 + 1 >= g) ? sc[0][1] : 0.f;
              sc[0][2] = (c0 >= g + 8) ? sc[0][2] : 0.f;     sc[0][3] = (c0 + 1 >= g + 8) ? sc[0][3] : 0.f;
              sc[1][0] = (c1 >= g) ? sc[1][0] : 0.f;         sc[1][1] = (c1 + 1 >= g) ? sc[1][1] : 0.f;
              sc[1][2] = (c1 >= g + 8) ? sc[1][2] : 0.f;     sc[1][3] = (c1 + 1 >= g + 8) ? sc[1][3] : 0.f;
            }
            uint32_t pa[4];
            pa[0] = pack_bf16x2(sc[0][0], sc[0][1]);
            pa[1] = pack_bf16x2(sc[0][2], sc[0][3]);
            pa[2] = pack_bf16x2(sc[1][0], sc[1][1]);
            pa[3] = pack_bf16x2(sc[1][2], sc[1][3]);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {           // B = G stored [K = query][N = 80]
              uint32_t bfr[4];
              ldsm4t(bfr, BKN_ADDR(sm.g, kk * 16, j * 8));
              mma16816(o[j], pa, bfr[0], bfr[1]);
              mma16816(o[j + 1], pa, bfr[2], bfr[3]);
            }
          }
        }
        bar_wait<BAR_DV>();                            // + phi(k) R, computed by query warp w4
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = sm.dvp[w4][j][lane];
          o[j][0] += t.x; o[j][1] += t.y; o[j][2] += t.z; o[j][3] += t.w;
        }
        bf16 (*stage)[LD64] = sm.xk[buf];              // this warp's rows (dk already copied out)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<uint32_t*>(&stage[r0 + g][8 * j + 2 * t4]) = pack_bf16x2(o[j][0], o[j][1]);
          *reinterpret_cast<uint32_t*>(&stage[r0 + g + 8][8 * j + 2 * t4]) = pack_bf16x2(o[j][2], o[j][3]);
        }
        __syncwarp();
        bf16* dst = dv + dbase + (int64_t)t0 * ld_d;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = lane + 32 * i, row = r0 + (idx >> 3), part = idx & 7;
          if (row < valid) *reinterpret_cast<uint4*>(dst + (int64_t)row * ld_d + part * 8) = *reinterpret_cast<const uint4*>(&stage[row][part * 8]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace favor2b

int emo_favor_bwd2_launch(const void* q, const void* k, const void* v, int64_t ld, const float* omega, const void* out,
                          const void* dout, int64_t ld_out, const float* den, const float* seg_states,
                          const float* seg_rstates, int nseg, int sc, int fwd_nseg, int ratio, void* dq, void* dk, void* dv,
                          int64_t ld_d, int B, int T_, int H, cudaStream_t s) {
  using namespace favor2b;
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(favor_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    configured = true;
  }
  favor_bwd2_kernel<<<B * H * nseg, NT_, sizeof(Smem), s>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, ld, omega, (const bf16*)out,
                                                           (const bf16*)dout, ld_out, den, seg_states, seg_rstates, nseg, sc,
                                                           fwd_nseg, ratio, (bf16*)dq, (bf16*)dk, (bf16*)dv, ld_d, T_, H);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
