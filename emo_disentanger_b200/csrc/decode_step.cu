// One-kernel decode step of the stage-2 Performer (reference stage2_accompaniment/inference.py:252-272 feeds the
// model one new token per iteration).
//
// The per-token work is ~76 MB of bf16 weights streamed once (L2 resident between tokens) and almost no arithmetic.
// As a chain of 62 small kernels -- or as one cooperative kernel with 61 GRID barriers, which was built and measured
// first -- the step costs ~3.5-4 us per dependent phase: a grid-wide rendezvous (kernel boundary or software barrier)
// is two L2 round trips whatever one does.  So each SEQUENCE gets one thread-block CLUSTER of 16 CTAs instead: the
// 61 phase boundaries become hardware cluster barriers (barrier.cluster, ~0.3 us), activations travel through L2
// between the CTAs of the cluster, and the 16 SMs stream the layer weights with two register buffers of 16-byte
// loads per lane in flight (the next phase's first buffer is issued BEFORE the barrier).  Sequences are independent
// clusters -- no cooperative launch, no grid-wide state, any batch size.
//
//   embed | for l: [LN2(l-1)] qkv GEMV | FAVOR+ recurrent step (one CTA per head) | out-proj GEMV + residual |
//           [LN1] FFN1 GEMV + ReLU | FFN2 GEMV + residual | [LN2(last)] logits GEMV
//
// Arithmetic is the per-column / per-row arithmetic of the multi-kernel path (gemm_skinny_nt_kernel, favor_step_kernel,
// embed_rows_kernel: same chunk order, same fma chains, same roundings), so the two paths agree bit for bit.
#include "common.cuh"

namespace {

constexpr int DS_THREADS = 512, DS_WARPS = 16, DS_CLUSTER = 16;
constexpr int D = 512, DF = 2048, DH = 8, DE = 64, DM = 128, DFV = 80;
constexpr float DS_EPS = 1e-6f;
constexpr float DS_HALF_LOG_M = 2.4260151319598084f;

struct DecodeArgs {
  const bf16* Wc;            // bf16 shadow of the flat parameter buffer
  const float* Wf;           // fp32 master flat buffer (biases, LayerNorm, embedding tables)
  const int64_t* offs;       // [L][12] element offsets: wqkv wo w1 w2 | bqkv bo b1 b2 | g1 be1 g2 be2
  int64_t off_tok, off_seg, off_outw, off_outb;
  const float* pe;           // [max_pos, 512] or null
  const float* omegas;       // [L, 64, 64]
  float* state;              // [L, B, H, 128, 80]
  const int64_t* tok; const int64_t* seg; int64_t* pos;
  bf16 *h, *qkv, *att, *s1, *y1, *hh, *s2;      // scratch activations [B, .]
  float* logits;             // [B, ldv]
  int L, B, V, ldv;
  float emb_scale;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// phase boundary: every global write of the cluster's threads before it is visible to all of them after it
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The weights of a warp's WHOLE phase: 16 x 16 bytes per lane = 8 output columns x K = 512, or 2 columns x K = 2048
// (256 warps per cluster: qkv 6 columns, out-proj 2, FFN1 8, FFN2 2, logits 2) -- all issued before the barrier.
struct WBuf { uint4 v[16]; float bias[8], res[8]; };    // + the columns' bias / residual values (same for every lane)

template <int K>
__device__ __forceinline__ void wload(WBuf& w, const bf16* W, int64_t N, int64_t n0, int64_t stride, const float* bias,
                                      const bf16* residual) {
  constexpr int VPL = (K >> 3) >> 5, NC = 16 / VPL;         // vectors per lane per column; columns per buffer
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int64_t n = n0 + j * stride;
    if (n < N) {
      const uint4* brow = reinterpret_cast<const uint4*>(W + n * (int64_t)K);
#pragma unroll
      for (int u = 0; u < VPL; ++u) w.v[j * VPL + u] = __ldg(brow + lane + 32 * u);
      w.bias[j] = bias ? __ldg(bias + n) : 0.f;
      w.res[j] = residual ? to_f(residual[n]) : 0.f;       // (plain load: written by a peer CTA one phase earlier)
    }
  }
}
// residual values of a buffer that was fetched before the barrier (its residual did not exist yet)
template <int K>
__device__ __forceinline__ void wload_res(WBuf& w, int64_t N, int64_t n0, int64_t stride, const bf16* residual) {
  constexpr int NC = 16 / ((K >> 3) >> 5);
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int64_t n = n0 + j * stride;
    if (n < N) w.res[j] = to_f(residual[n]);
  }
}

// the sequence's activation row [K] (global, bf16) -> shared; optional LayerNorm over K == 512 (fp32 statistics, bf16
// result); CTA rank 0 also stores the normalised row to ln_out.  Every CTA keeps its own copy (2-4 KB).
__device__ __forceinline__ void stage_row(bf16* As, const bf16* A, int K, const float* gamma, const float* beta, bf16* ln_out,
                                          bool writer) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kv = K >> 3;
  if ((int)threadIdx.x < kv) reinterpret_cast<uint4*>(As)[threadIdx.x] = reinterpret_cast<const uint4*>(A)[threadIdx.x];
  __syncthreads();
  if (gamma) {
    if (warp == 0) {
      uint4* row = reinterpret_cast<uint4*>(As);
      float x[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 t = row[lane + 32 * h];
        unpack_bf16x2(t.x, x[8 * h], x[8 * h + 1]); unpack_bf16x2(t.y, x[8 * h + 2], x[8 * h + 3]);
        unpack_bf16x2(t.z, x[8 * h + 4], x[8 * h + 5]); unpack_bf16x2(t.w, x[8 * h + 6], x[8 * h + 7]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) sum += x[j];
      const float mu = warp_sum(sum) * (1.f / 512.f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { float d = x[j] - mu; q += d * d; }
      const float rs = rsqrtf(warp_sum(q) * (1.f / 512.f) + 1e-5f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = (lane + 32 * h) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
        const float gq[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (x[8 * h + j] - mu) * rs * gq[j] + bq[j];
        uint4 t;
        t.x = pack_bf16x2(y[0], y[1]); t.y = pack_bf16x2(y[2], y[3]); t.z = pack_bf16x2(y[4], y[5]); t.w = pack_bf16x2(y[6], y[7]);
        row[lane + 32 * h] = t;
        if (writer && ln_out) *reinterpret_cast<uint4*>(ln_out + c0) = t;
      }
    }
    __syncthreads();
  }
}

// the columns of one register buffer: out[n] = act(As . W[n] + bias[n]) + residual[n]
template <typename TOut, bool RELU, int K>
__device__ __forceinline__ void wcompute(const WBuf& w, const bf16* As, int64_t N, int64_t n0, int64_t stride, TOut* out) {
  constexpr int VPL = (K >> 3) >> 5, NC = 16 / VPL;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int64_t n = n0 + j * stride;
    if (n < N) {
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < VPL; ++u) {          // chunk c = lane + 32 u: the order of gemm_skinny_nt_kernel
        const int c = lane + 32 * u;
        float b[8], a[8];
        const uint4 bw = w.v[j * VPL + u];
        unpack_bf16x2(bw.x, b[0], b[1]); unpack_bf16x2(bw.y, b[2], b[3]);
        unpack_bf16x2(bw.z, b[4], b[5]); unpack_bf16x2(bw.w, b[6], b[7]);
        const uint4 aw = reinterpret_cast<const uint4*>(As)[c];
        unpack_bf16x2(aw.x, a[0], a[1]); unpack_bf16x2(aw.y, a[2], a[3]);
        unpack_bf16x2(aw.z, a[4], a[5]); unpack_bf16x2(aw.w, a[6], a[7]);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc = fmaf(a[jj], b[jj], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        float v = acc + w.bias[j];
        if (RELU) v = fmaxf(v, 0.f);
        out[n] = from_f<TOut>(v + w.res[j]);
      }
    }
  }
}

// One GEMV phase for this warp: its columns are n = cw + k * CW (cw = cluster-wide warp id), taken NC at a time.
// `first` holds them all (issued before the barrier that opened the phase).
template <typename TOut, bool RELU, int K>
__device__ __forceinline__ void gemv_phase(WBuf& first, const bf16* As, const bf16* W, int64_t N, const float* bias,
                                           const bf16* residual, TOut* out, int64_t cw, int64_t CW) {
  constexpr int NC = 16 / ((K >> 3) >> 5);
  int64_t n0 = cw;
  if (n0 >= N) return;
  if (residual) wload_res<K>(first, N, n0, CW, residual);
  wcompute<TOut, RELU, K>(first, As, N, n0, CW, out);
  for (n0 += NC * CW; n0 < N; n0 += NC * CW) {          // (only with fewer warps per cluster than columns / NC)
    wload<K>(first, W, N, n0, CW, bias, residual);
    wcompute<TOut, RELU, K>(first, As, N, n0, CW, out);
  }
}

// recurrent FAVOR+ step of one (sequence, head): same arithmetic as favor_step_kernel (favor_kernels.cuh)
__device__ __forceinline__ void favor_step_cta(float* sm, int bh, const bf16* qkv, const float* omega, float* state, bf16* att) {
  constexpr int CG_ = 17, RL = 15, NR = (DM + RL - 1) / RL;
  float* xq = sm; float* xk = xq + DE; float* pq = xk + DE; float* pk = pq + DM;
  float* vv = pk + DM;                    // [68], 16-byte aligned (offset 384 floats)
  float* red = vv + CG_ * 4;              // [15][68]
  float* om_s = red + RL * CG_ * 4;       // [4096], offset 384 + 68 + 1020 = 1472 floats (16-byte aligned)
  const int b = bh / DH, h = bh % DH, tid = threadIdx.x;
  const float s = 0.35355339059327373f;
  if (tid < 256) {        // (512 threads per CTA; this routine is written for 256 workers, all take the barriers)
    float4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = __ldg(reinterpret_cast<const float4*>(omega) + tid + 256 * u);
#pragma unroll
    for (int u = 0; u < 4; ++u) reinterpret_cast<float4*>(om_s)[tid + 256 * u] = t[u];
  }
  float4 st[NR];
  if (tid < CG_ * RL) {
    const float* sb0 = state + (int64_t)bh * DM * DFV + (tid % CG_) * 4;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int f = tid / CG_ + i * RL;
      if (f < DM) st[i] = *reinterpret_cast<const float4*>(sb0 + f * DFV);
    }
  }
  const bf16* q = qkv + (int64_t)b * 3 * D + h * DE;
  if (tid < DE) {
    xq[tid] = to_f(q[tid]) * s;
    xk[tid] = to_f(q[D + tid]) * s;
    vv[tid] = to_f(q[2 * D + tid]);
  } else if (tid < CG_ * 4) {
    vv[tid] = (tid == DE) ? 1.f : 0.f;
  }
  __syncthreads();
  if (tid < 2 * DE) {
    const float* x = (tid < DE) ? xq : xk;
    int f = tid & (DE - 1);
    float u = 0.f, n2 = 0.f;
#pragma unroll 8
    for (int e = 0; e < DE; ++e) { u = fmaf(x[e], om_s[e * DE + f], u); n2 = fmaf(x[e], x[e], n2); }
    float o = 0.5f * n2 + DS_HALF_LOG_M;
    float* p = (tid < DE) ? pq : pk;
    p[f] = expf(u - o);
    p[DE + f] = expf(-u - o);
  }
  __syncthreads();
  if (tid < CG_ * RL) {
    const int cgi = tid % CG_, rl = tid / CG_;
    const float4 v4 = *reinterpret_cast<const float4*>(&vv[cgi * 4]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float* sbase = state + (int64_t)bh * DM * DFV + cgi * 4;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int f = rl + i * RL;
      if (f >= DM) break;
      float4* sp = reinterpret_cast<float4*>(sbase + f * DFV);
      float4 sv = st[i];
      const float a = pk[f], c = pq[f];
      sv.x = fmaf(a, v4.x, sv.x); sv.y = fmaf(a, v4.y, sv.y); sv.z = fmaf(a, v4.z, sv.z); sv.w = fmaf(a, v4.w, sv.w);
      *sp = sv;
      acc.x = fmaf(c, sv.x, acc.x); acc.y = fmaf(c, sv.y, acc.y); acc.z = fmaf(c, sv.z, acc.z); acc.w = fmaf(c, sv.w, acc.w);
    }
    *reinterpret_cast<float4*>(&red[rl * CG_ * 4 + cgi * 4]) = acc;
  }
  __syncthreads();
  if (tid <= DE) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < RL; ++r) t += red[r * CG_ * 4 + tid];
    red[tid] = t;
  }
  __syncthreads();
  if (tid < DE) att[(int64_t)b * D + h * DE + tid] = from_f<bf16>(red[tid] / (red[DE] + DS_EPS));
  __syncthreads();
}


__global__ void __launch_bounds__(DS_THREADS, 1) performer_decode_step_kernel(const __grid_constant__ DecodeArgs a) {
  __shared__ __align__(16) unsigned char ds_smem[24 * 1024];     // activation row (<= 4 KB) / FAVOR+ step scratch (22 KB)
  bf16* As = reinterpret_cast<bf16*>(ds_smem);
  const int warp = threadIdx.x >> 5;
  const int CS = (int)cluster_size(), rank = (int)cluster_rank();
  const int b = blockIdx.x / CS;                                  // one cluster per sequence
  const int64_t cw = (int64_t)rank * DS_WARPS + warp, CW = (int64_t)CS * DS_WARPS;
  const bool writer = rank == 0;
  bf16 *h = a.h + (int64_t)b * D, *qkv = a.qkv + (int64_t)b * 3 * D, *att = a.att + (int64_t)b * D, *s1 = a.s1 + (int64_t)b * D,
       *y1 = a.y1 + (int64_t)b * D, *hh = a.hh + (int64_t)b * DF, *s2 = a.s2 + (int64_t)b * D;
  WBuf wb;

  // ---- embedding row ----
  wload<D>(wb, a.Wc + a.offs[0], 3 * D, cw, CW, a.Wf + a.offs[4], nullptr);
  if (writer) {
    const float* e = a.Wf + a.off_tok + a.tok[b] * D;
    const float* sg = (a.off_seg >= 0) ? a.Wf + a.off_seg + a.seg[b] * D : nullptr;
    const int64_t pr = a.pe ? a.pos[b] : 0;
    const float* p = a.pe ? a.pe + pr * D : nullptr;
    for (int c = threadIdx.x; c < D; c += DS_THREADS) {
      float v = e[c] * a.emb_scale;
      if (sg) v += sg[c] * a.emb_scale;
      if (p) v += p[c];
      h[c] = from_f<bf16>(v);
    }
    if (a.pe) {
      __syncthreads();
      if (threadIdx.x == 0) a.pos[b] = pr + 1;
    }
  }
  cluster_barrier();

  for (int l = 0; l < a.L; ++l) {
    const int64_t* o = a.offs + l * 12;
    const int64_t* op = a.offs + (l > 0 ? l - 1 : 0) * 12;
    // ---- A: qkv = [LN2 of the previous layer](s2 or h) . Wqkv^T + b ----
    stage_row(As, l == 0 ? h : s2, D, l > 0 ? a.Wf + op[10] : nullptr, l > 0 ? a.Wf + op[11] : nullptr, h, writer);
    gemv_phase<bf16, false, D>(wb, As, a.Wc + o[0], 3 * D, a.Wf + o[4], nullptr, qkv, cw, CW);
    cluster_barrier();
    // ---- B: FAVOR+ recurrent step, one CTA per head (no weights are held across it: registers) ----
    for (int hd = rank; hd < DH; hd += CS)
      favor_step_cta(reinterpret_cast<float*>(ds_smem), b * DH + hd, a.qkv, a.omegas + (int64_t)l * DE * DE,
                     a.state + (int64_t)l * a.B * DH * DM * DFV, a.att);
    wload<D>(wb, a.Wc + o[1], D, cw, CW, a.Wf + o[5], nullptr);
    cluster_barrier();
    // ---- C: s1 = att . Wo^T + bo + h ----
    stage_row(As, att, D, nullptr, nullptr, nullptr, false);
    gemv_phase<bf16, false, D>(wb, As, a.Wc + o[1], D, a.Wf + o[5], h, s1, cw, CW);
    wload<D>(wb, a.Wc + o[2], DF, cw, CW, a.Wf + o[6], nullptr);
    cluster_barrier();
    // ---- D: hh = relu(LN1(s1) . W1^T + b1);  y1 = LN1(s1) ----
    stage_row(As, s1, D, a.Wf + o[8], a.Wf + o[9], y1, writer);
    gemv_phase<bf16, true, D>(wb, As, a.Wc + o[2], DF, a.Wf + o[6], nullptr, hh, cw, CW);
    wload<DF>(wb, a.Wc + o[3], D, cw, CW, a.Wf + o[7], nullptr);
    cluster_barrier();
    // ---- E: s2 = hh . W2^T + b2 + y1 ----
    stage_row(As, hh, DF, nullptr, nullptr, nullptr, false);
    gemv_phase<bf16, false, DF>(wb, As, a.Wc + o[3], D, a.Wf + o[7], y1, s2, cw, CW);
    if (l + 1 < a.L) wload<D>(wb, a.Wc + a.offs[(l + 1) * 12], 3 * D, cw, CW, a.Wf + a.offs[(l + 1) * 12 + 4], nullptr);
    else wload<D>(wb, a.Wc + a.off_outw, a.V, cw, CW, a.Wf + a.off_outb, nullptr);
    cluster_barrier();
  }
  // ---- logits = LN2(last)(s2) . Wout^T + bout (fp32) ----
  const int64_t* ol = a.offs + (a.L - 1) * 12;
  stage_row(As, s2, D, a.Wf + ol[10], a.Wf + ol[11], nullptr, false);
  gemv_phase<float, false, D>(wb, As, a.Wc + a.off_outw, a.V, a.Wf + a.off_outb, nullptr, a.logits + (int64_t)b * a.ldv, cw, CW);
}

}  // namespace

// ---- C ABI ----------------------------------------------------------------------------------
extern "C" int emo_performer_decode_step(const void* w_bf16, const float* w_f32, const int64_t* layer_offs,
                                         int64_t off_tok, int64_t off_seg, int64_t off_outw, int64_t off_outb,
                                         const float* pe, const float* omegas, float* state, const int64_t* tok,
                                         const int64_t* seg, int64_t* pos, void* scratch, float* logits, int n_layer,
                                         int batch, int n_token, int ld_logits, float emb_scale, void* stream) {
  EMO_REQUIRE(batch >= 1 && n_layer >= 1 && n_token >= 1, "emo_performer_decode_step: bad sizes");
  DecodeArgs a;
  a.Wc = (const bf16*)w_bf16; a.Wf = w_f32; a.offs = layer_offs;
  a.off_tok = off_tok; a.off_seg = off_seg; a.off_outw = off_outw; a.off_outb = off_outb;
  a.pe = pe; a.omegas = omegas; a.state = state; a.tok = tok; a.seg = seg; a.pos = pos;
  bf16* s = (bf16*)scratch;            // [B] x (512 h | 1536 qkv | 512 att | 512 s1 | 512 y1 | 2048 hh | 512 s2)
  a.h = s; s += (int64_t)batch * D;
  a.qkv = s; s += (int64_t)batch * 3 * D;
  a.att = s; s += (int64_t)batch * D;
  a.s1 = s; s += (int64_t)batch * D;
  a.y1 = s; s += (int64_t)batch * D;
  a.hh = s; s += (int64_t)batch * DF;
  a.s2 = s;
  a.logits = logits; a.L = n_layer; a.B = batch; a.V = n_token; a.ldv = ld_logits; a.emb_scale = emb_scale;
  static int cluster = 0;               // 16 CTAs per cluster where the device can place them, else 8
  if (cluster == 0) {
    cluster = 8;
    if (cudaFuncSetAttribute(performer_decode_step_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(DS_CLUSTER); q.blockDim = dim3(DS_THREADS);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = DS_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, performer_decode_step_kernel, &q) == cudaSuccess && nc >= 1) cluster = DS_CLUSTER;
    }
    (void)cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * cluster);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  EMO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, performer_decode_step_kernel, a));
  return EMO_OK;
}
