// K9 on the 5th-gen tensor cores: the stage-1 relative-position attention (stage1_compose/model/optimus_txl_decoder.py:
// 305-387, RelPartialLearnableMultiHeadAttn) built from the tcgen05 attention kernels of attn_tc.cu + the tcgen05 GEMM.
//
//   score(i, j) = scale * ( (q_i + r_w_bias) . k_j  +  (q_i + r_r_bias) . r_{p(i,j)} ),   p(i, j) = Tq - 1 - i + j
//
// The position term is a Toeplitz-shifted read of G = (q + r_r_bias) r^T (the reference's _rel_shift): row i of the
// score tile needs G[i][Tq - 1 - i + j], a per-ROW column shift.  A tcgen05.ld addresses the same columns for all 32
// lanes of a warp and an MMA cannot produce the shifted tile directly (no operand depends on i alone or j alone), so
// the shift is done by ADDRESSING in HBM instead of in registers:
//   forward    qw = q + r_w_bias                                                 (one element-wise pass)
//              BD (queries x keys, blocked planes) = shifted (q + r_r_bias) r^T  (rel_scores_shift_kernel: per 64 x 64
//                                                                                 tile one mma.sync product over the 127
//                                                                                 reachable rows of r, stored shifted)
//              attention forward with BD as an additive score tile               (attn_fwd_tc_kernel, rel mode)
//   backward   BD^T (keys x queries) re-made (cheaper than keeping 0.5 GB per layer alive), attention backward adds
//              BD^T to K Q^T and writes dS^T next to it; dG = un-shifted, transposed dS^T; then three GEMM families:
//              dqv_h = dG_h r_h,  dr_h += dG_h^T qv_h,  and dq = dqw (attention) + dqv; bias gradients = column sums.
// Position scores travel as bf16 (like every activation of the bf16 path); everything is fp32 inside the kernels.
// Shapes that do not fit (Tq < 64, Tk % 64, Tq % 32, fp32) stay on attn.cu's mma.sync kernels.
#include "block_gemm.cuh"

struct AttnTcRel { const void* bias; const void* biasT; void* dbiasT; int rel; };
// Layout of the bf16 score planes handed to / received from the attention kernels (BD: rows = queries, columns = keys;
// BD^T and dBD^T: rows = keys, columns = queries): per (batch x head) a grid of 128 x 128 blocks, and inside a block the
// 16-byte chunk (8 columns) c of row r sits at ((r / 32 * 16 + c) * 32 + r % 32) * 16 bytes -- the 32 rows a warp of the
// attention kernels owns are contiguous for every chunk, so each of its load / store instructions touches 4 lines.  With
// row-major planes (one row per lane) every instruction touched 32 lines and the kernels were bound by the load / store
// unit in rel mode.
__host__ __device__ __forceinline__ int64_t rel_plane_elems(int rows, int cols) {
  return (int64_t)((rows + 127) / 128) * ((cols + 127) / 128) * 16384;
}
__device__ __forceinline__ int64_t rel_blocked_off(int row, int col, int cols) {
  const int nct = (cols + 127) >> 7, rr = row & 127, cc = col & 127;
  return ((int64_t)(row >> 7) * nct + (col >> 7)) * 16384 + (((rr >> 5) * 16 + (cc >> 3)) * 32 + (rr & 31)) * 8 + (cc & 7);
}
int emo_attn_fwd_tc_launch_ex(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out, int64_t ld_o,
                              float* lse, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, const AttnTcRel* rel,
                              cudaStream_t s);
int64_t emo_attn_bwd_tc_ws_floats(int B, int Tq, int H);
int emo_attn_tc_configure_pool();
int emo_attn_bwd_tc_core(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* out, const void* dout,
                         int64_t ld_o, const float* lse, float* ws, void* dk, void* dv, int64_t ld_dkv, int B, int Tq, int Tk, int H,
                         float scale, float drop_p, uint64_t seed, const AttnTcRel* rel, cudaStream_t s);
int emo_attn_bwd_tc_convert(const float* dqacc, const void* add, void* dq, int64_t ld_dq, int B, int Tq, int H, cudaStream_t s);

namespace {
constexpr int RE = 64;

// qw = q + r_w_bias, qv = q + r_r_bias: dense [rows][d] bf16 copies of the strided q (8 elements per thread)
__global__ void rel_prep_kernel(const bf16* __restrict__ q, int64_t ld_q, const float* __restrict__ rwb, const float* __restrict__ rrb,
                                bf16* __restrict__ qw, bf16* __restrict__ qv, int64_t rows, int d) {
  const int vpr = d / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * vpr) return;
  const int64_t row = idx / vpr;
  const int c = (int)(idx % vpr) * 8;
  Vec<bf16> x, a, b;
  x.load(q + row * ld_q + c);
#pragma unroll
  for (int j = 0; j < 8; ++j) { a.v[j] = x.v[j] + rwb[c + j]; b.v[j] = x.v[j] + rrb[c + j]; }
  a.store(qw + row * d + c);
  b.store(qv + row * d + c);
}

// dG[h][b*Tq + i][t] = dBDT[bh][j][i] with j = t - (Tq - 1) + i when 0 <= j <= min(i + off, Tk - 1), else 0.
// Output tile 64 (i) x 64 (t); its sources are 127 rows j of 64 consecutive i.
__global__ void __launch_bounds__(256) rel_unshift_kernel(const bf16* __restrict__ dBDT, bf16* __restrict__ dG, int B, int H, int Tq, int Tk) {
  __shared__ unsigned short patch[127][66];
  const int i0 = blockIdx.x * 64, t0 = blockIdx.y * 64;
  const int64_t bh = blockIdx.z;
  const int b = (int)(bh / H), h = (int)(bh % H);
  const int off = Tk - Tq;
  const int tx = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int jmin = t0 - (Tq - 1) + i0;
  // whole 16-byte chunks of the blocked plane: lane = (row & 3, chunk of 8 queries) -- the same chunk of 4 consecutive key
  // rows is 64 contiguous bytes there, so every sector that is fetched is used (4-byte loads per query pair used half)
  const bf16* plane = dBDT + bh * rel_plane_elems(Tk, Tq);
  const int rsub = tx & 3, c8 = tx >> 2;
  for (int g4 = w; g4 < 32; g4 += 8) {
    const int jr = 4 * g4 + rsub, j = jmin + jr, ic = i0 + 8 * c8;
    if (jr >= 127) continue;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (j >= 0 && j < Tk && ic < Tq && j <= ic + 7 + off)          // (Tq % 32 == 0: a chunk is all-in or all-out)
      v = __ldg(reinterpret_cast<const uint4*>(plane + rel_blocked_off(j, ic, Tq)));
    uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {                                  // query i sees key j iff j <= i + off
      const int i = ic + 2 * e;
      if (j > i + 1 + off) wv[e] = 0u;
      else if (j == i + 1 + off) wv[e] &= 0xffff0000u;
      *reinterpret_cast<uint32_t*>(&patch[jr][8 * c8 + 2 * e]) = wv[e];
    }
  }
  __syncthreads();
  if (t0 + 2 * tx >= Tk) return;                 // Tk is a multiple of 64
#pragma unroll
  for (int ii = w; ii < 64; ii += 8) {
    const int i = i0 + ii;
    if (i < Tq) {
      const uint32_t v = (uint32_t)patch[2 * tx + ii][ii] | ((uint32_t)patch[2 * tx + 1 + ii][ii] << 16);
      *reinterpret_cast<uint32_t*>(dG + ((int64_t)h * B * Tq + (int64_t)b * Tq + i) * Tk + t0 + 2 * tx) = v;
    }
  }
}

// Position scores of one 64 x 64 (query, key) tile straight into their shifted place: G = (q + r_r_bias) . band^T over the
// 127 rows of r the tile can reach (mma.sync, 64 x 128 x 64), then BD[ii][jj] = G[ii][63 - ii + jj] read back from shared
// memory at a row-dependent offset and stored as 16-byte vectors -- row-major (forward) or transposed (backward).
// Replaces eight K = 64 GEMMs that wrote G to HBM plus a shifting copy that read it back.  Tiles without a visible pair
// are skipped: the attention kernels mask those positions before they look at the value.
struct RelScoreSmem {
  bf16 qv[64][bg_ld<bf16>(RE)];
  bf16 rb[128][bg_ld<bf16>(RE)];
  unsigned short g[64][134];      // G as bf16 bits (rounded once, here: what the planes store anyway): 45 KB per CTA -> 5 CTAs per SM
};
template <bool TRANS>
__global__ void __launch_bounds__(BG_THREADS) rel_scores_shift_kernel(const bf16* __restrict__ q, int64_t ld_q, const float* __restrict__ rrb,
                                                                      const bf16* __restrict__ r, int64_t ld_r, bf16* __restrict__ outp,
                                                                      int H, int Tq, int Tk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RelScoreSmem& sm = *reinterpret_cast<RelScoreSmem*>(smem_raw);
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64, off = Tk - Tq;
  if (j0 > i0 + 63 + off) return;                            // block-uniform: nothing of this tile is visible
  const int64_t bh = blockIdx.z;
  const int b = (int)(bh / H), h = (int)(bh % H);
  const int tid = threadIdx.x;
  for (int t = tid; t < 64 * 8; t += BG_THREADS) {           // q rows + r_r_bias -> bf16
    const int rr = t >> 3, part = t & 7, i = i0 + rr;
    Vec<bf16> x;
    if (i < Tq) x.load(q + ((int64_t)b * Tq + i) * ld_q + h * RE + part * 8);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x.v[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm.qv[rr][part * 8 + e] = from_f<bf16>(x.v[e] + rrb[h * RE + part * 8 + e]);
  }
  const int pbase = Tq - i0 - 64 + j0;                       // r row of (ii, jj) is pbase + 63 - ii + jj
  for (int t = tid; t < 128 * 8; t += BG_THREADS) {
    const int rr = t >> 3, part = t & 7, pr = pbase + rr;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (pr >= 0 && pr < Tk) v = *reinterpret_cast<const uint4*>(r + (int64_t)pr * ld_r + h * RE + part * 8);
    *reinterpret_cast<uint4*>(&sm.rb[rr][part * 8]) = v;
  }
  __syncthreads();
  {
    BlockGemm<64, 128, bf16> gg;
    gg.clear();
    gg.template mma<true, true>(&sm.qv[0][0], bg_ld<bf16>(RE), &sm.rb[0][0], bg_ld<bf16>(RE), RE);
    gg.foreach ([&](int r_, int c_, float& x) { sm.g[r_][c_] = __bfloat16_as_ushort(__float2bfloat16_rn(x)); });
  }
  __syncthreads();
  const int row = tid >> 2, chunk = tid & 3;                 // 64 rows x 4 chunks of 16 elements
  uint32_t w[8];
  if (!TRANS) {
    const int i = i0 + row;
    if (i >= Tq) return;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int jj = 16 * chunk + 2 * e;
      const uint32_t v0 = (j0 + jj <= i + off) ? sm.g[row][63 - row + jj] : 0u;
      const uint32_t v1 = (j0 + jj + 1 <= i + off) ? sm.g[row][63 - row + jj + 1] : 0u;
      w[e] = v0 | (v1 << 16);
    }
    bf16* plane = outp + bh * rel_plane_elems(Tq, Tk);
    *reinterpret_cast<uint4*>(plane + rel_blocked_off(i, j0 + 16 * chunk, Tk)) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(plane + rel_blocked_off(i, j0 + 16 * chunk + 8, Tk)) = make_uint4(w[4], w[5], w[6], w[7]);
  } else {
    const int j = j0 + row;
    if (j >= Tk || i0 + 16 * chunk >= Tq) return;            // (Tq is a multiple of 32: 16-element chunks are all-in or all-out)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ii = 16 * chunk + 2 * e;
      const uint32_t v0 = (j <= i0 + ii + off) ? sm.g[ii][63 - ii + row] : 0u;
      const uint32_t v1 = (j <= i0 + ii + 1 + off) ? sm.g[ii + 1][62 - ii + row] : 0u;
      w[e] = v0 | (v1 << 16);
    }
    bf16* plane = outp + bh * rel_plane_elems(Tk, Tq);
    *reinterpret_cast<uint4*>(plane + rel_blocked_off(j, i0 + 16 * chunk, Tq)) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(plane + rel_blocked_off(j, i0 + 16 * chunk + 8, Tq)) = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
template <bool TRANS>
int rel_scores_shift(const void* q, int64_t ld_q, const float* rrb, const void* r, int64_t ld_r, bf16* outp, int B, int Tq, int Tk, int H,
                     cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    EMO_CHECK_CUDA(cudaFuncSetAttribute(rel_scores_shift_kernel<TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RelScoreSmem)));
    configured = true;
  }
  dim3 grid((Tq + 63) / 64, (Tk + 63) / 64, B * H);
  rel_scores_shift_kernel<TRANS><<<grid, BG_THREADS, sizeof(RelScoreSmem), s>>>((const bf16*)q, ld_q, rrb, (const bf16*)r, ld_r, outp, H, Tq, Tk);
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

}  // namespace

bool emo_relattn_tc_ok(int B, int Tq, int Tk, int H) {
  // (a single short sequence is launch-bound: ~45 launches here against 3 of the mma.sync kernels)
  return Tq >= 64 && (Tk % 64) == 0 && (Tq % 32) == 0 && (int64_t)B * Tq >= 2048 && (int64_t)B * H * Tq * (int64_t)Tk < (1ll << 40);
}

int emo_relattn_fwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* r, int64_t ld_r,
                              const float* r_w_bias, const float* r_r_bias, void* out, int64_t ld_out, float* lse, int B, int Tq, int Tk,
                              int H, float scale, float drop_p, uint64_t seed, cudaStream_t s) {
  int rc = emo_attn_tc_configure_pool();
  if (rc) return rc;
  const int d = H * RE;
  const int64_t rows = (int64_t)B * Tq, nP = (int64_t)B * H * rel_plane_elems(Tq, Tk);
  const bool padded = (Tq % 128) != 0 || (Tk % 128) != 0;      // block rows / columns nobody writes must still be finite
  bf16* ws = nullptr;
  EMO_CHECK_CUDA(cudaMallocAsync((void**)&ws, (size_t)(2 * rows * d + nP + 512) * sizeof(bf16), s));
  bf16 *qw = ws, *qv = qw + rows * d, *BD = qv + rows * d;
  do {
    if (padded && cudaMemsetAsync(BD, 0, (size_t)nP * sizeof(bf16), s) != cudaSuccess) break;
    rel_prep_kernel<<<(unsigned)((rows * (d / 8) + 255) / 256), 256, 0, s>>>((const bf16*)q, ld_q, r_w_bias, r_r_bias, qw, qv, rows, d);
    if ((rc = rel_scores_shift<false>(q, ld_q, r_r_bias, r, ld_r, BD, B, Tq, Tk, H, s))) break;
    AttnTcRel ex = {BD, nullptr, nullptr, 1};
    rc = emo_attn_fwd_tc_launch_ex(qw, k, v, d, ld_kv, out, ld_out, lse, B, Tq, Tk, H, scale, drop_p, seed, &ex, s);
  } while (0);
  cudaError_t e1 = cudaGetLastError(), e2 = cudaFreeAsync(ws, s);
  if (rc) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    emo_set_error("emo_relattn_fwd (tcgen05): CUDA error %d (%s)", (int)(e1 != cudaSuccess ? e1 : e2), cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return EMO_ERR_CUDA;
  }
  return EMO_OK;
}

int emo_relattn_bwd_tc_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, const void* r, int64_t ld_r,
                              const float* r_w_bias, const float* r_r_bias, const void* out, const void* dout, int64_t ld_out,
                              const float* lse, void* dq, void* dk, void* dv, int64_t ld_dq, int64_t ld_dkv, float* dr, float* d_rw,
                              float* d_rr, int B, int Tq, int Tk, int H, float scale, float drop_p, uint64_t seed, cudaStream_t s) {
  int rc = emo_attn_tc_configure_pool();
  if (rc) return rc;
  const int d = H * RE;
  const int64_t rows = (int64_t)B * Tq, nG = (int64_t)H * rows * Tk;
  const int64_t nf = emo_attn_bwd_tc_ws_floats(B, Tq, H);
  bf16* ws = nullptr;
  float* wf = nullptr;
  const int64_t nP = (int64_t)B * H * rel_plane_elems(Tk, Tq);
  const bool padded = (Tq % 128) != 0 || (Tk % 128) != 0;
  EMO_CHECK_CUDA(cudaMallocAsync((void**)&ws, (size_t)(3 * rows * d + nG + 2 * nP + 512) * sizeof(bf16), s));
  if (cudaMallocAsync((void**)&wf, (size_t)nf * sizeof(float), s) != cudaSuccess) {
    cudaFreeAsync(ws, s);
    emo_set_error("emo_relattn_bwd (tcgen05): workspace allocation failed");
    return EMO_ERR_CUDA;
  }
  bf16 *qw = ws, *qv = qw + rows * d, *dqv = qv + rows * d, *dG = dqv + rows * d, *BDT = dG + nG, *dBDT = BDT + nP;
  do {
    if (padded && cudaMemsetAsync(BDT, 0, (size_t)nP * sizeof(bf16), s) != cudaSuccess) break;
    rel_prep_kernel<<<(unsigned)((rows * (d / 8) + 255) / 256), 256, 0, s>>>((const bf16*)q, ld_q, r_w_bias, r_r_bias, qw, qv, rows, d);
    if ((rc = rel_scores_shift<true>(q, ld_q, r_r_bias, r, ld_r, BDT, B, Tq, Tk, H, s))) break;
    dim3 tgrid((Tq + 63) / 64, (Tk + 63) / 64, B * H);
    AttnTcRel ex = {nullptr, BDT, dBDT, 1};
    if ((rc = emo_attn_bwd_tc_core(qw, k, v, d, ld_kv, out, dout, ld_out, lse, wf, dk, dv, ld_dkv, B, Tq, Tk, H, scale, drop_p, seed, &ex, s))) break;
    rel_unshift_kernel<<<tgrid, 256, 0, s>>>(dBDT, dG, B, H, Tq, Tk);
    for (int h = 0; h < H && !rc; ++h) {
      // dqv_h [B*Tq, 64] = dG_h [B*Tq, Tk] . r_h [Tk, 64]
      rc = emo_gemm(EMO_GEMM_NN, rows, RE, Tk, dG + (int64_t)h * rows * Tk, Tk, (const bf16*)r + h * RE, ld_r, dqv + h * RE, d, EMO_BF16,
                    EMO_BF16, nullptr, s);
      if (rc) break;
      // dr_h [Tk, 64] (fp32, row stride H*64) += dG_h^T . qv_h
      emo_epilogue ep;
      memset(&ep, 0, sizeof(ep));
      ep.alpha = 1.f;
      ep.accumulate = 1;
      rc = emo_gemm(EMO_GEMM_TN, Tk, RE, rows, dG + (int64_t)h * rows * Tk, Tk, qv + h * RE, d, dr + h * RE, (int64_t)H * RE, EMO_BF16, EMO_F32,
                    &ep, s);
    }
    if (rc) break;
    // bias gradients: r_w_bias sees the content part of dq, r_r_bias the position part
    if ((rc = emo_colsum(wf, d, rows, d, d_rw, EMO_F32, s))) break;
    if ((rc = emo_colsum(dqv, d, rows, d, d_rr, EMO_BF16, s))) break;
    rc = emo_attn_bwd_tc_convert(wf, dqv, dq, ld_dq, B, Tq, H, s);
  } while (0);
  cudaError_t e1 = cudaGetLastError(), e2 = cudaFreeAsync(ws, s), e3 = cudaFreeAsync(wf, s);
  if (rc) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    emo_set_error("emo_relattn_bwd (tcgen05): CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
    return EMO_ERR_CUDA;
  }
  return EMO_OK;
}
