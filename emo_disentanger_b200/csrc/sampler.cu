// K11: temperature + nucleus (top-p) sampling on device, one CTA per row (V <= 1024).
// Mirrors stage2_accompaniment/inference.py:71-100 / stage1_compose/inference_utils.py:14-41:
// softmax(l/t) -> sort descending -> cumulative sum -> cut at the SECOND index above top_p (the
// first crossing token is kept) -> renormalise -> inverse-CDF draw with the caller's uniform.
// greedy mode = argmax with numpy's lowest-index tie break (bit-exact decode mode).
#include "common.cuh"

constexpr int SMP_N = 1024;
constexpr int SMP_THREADS = 256;

// The draw for ONE row by one CTA of SMP_THREADS threads: `row` = its V logits (global or shared memory), r = its index
// into u / out / status / banned / t_rows.  Shared by sample_kernel and the fused logits + sampler kernel below.
__device__ __forceinline__ void sample_row(const float* row, int r, int V, float inv_t, float top_p,
                                           const float* __restrict__ u, int greedy, int64_t* __restrict__ out,
                                           int32_t* __restrict__ status, const uint8_t* __restrict__ banned,
                                           const float* __restrict__ t_rows) {
  __shared__ float key[SMP_N];
  __shared__ int idx[SMP_N];
  __shared__ float red[SMP_THREADS / 32];
  __shared__ int redi[SMP_THREADS / 32];
  __shared__ float wsum[SMP_THREADS / 32];
  __shared__ int s_first, s_choice;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (t_rows && !greedy) inv_t = 1.f / t_rows[r];      // one temperature per row (a batch of quadrants)

  // ---- max / argmax ----
  float mx = -INFINITY;
  int am = 0x7fffffff;
  for (int c = tid; c < V; c += SMP_THREADS) {
    float v = row[c];
    if (v > mx || (v == mx && c < am)) { mx = v; am = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (ov > mx || (ov == mx && oa < am)) { mx = ov; am = oa; }
  }
  if (lane == 0) { red[w] = mx; redi[w] = am; }
  __syncthreads();
  mx = red[0]; am = redi[0];
#pragma unroll
  for (int i = 1; i < SMP_THREADS / 32; ++i)
    if (red[i] > mx || (red[i] == mx && redi[i] < am)) { mx = red[i]; am = redi[i]; }
  if (greedy) {
    if (tid == 0) { out[r] = am; if (status) status[r] = 0; }
    return;
  }
  __syncthreads();

  // ---- softmax(l / t) ----
  float se = 0.f;
  for (int c = tid; c < SMP_N; c += SMP_THREADS) {
    float e = (c < V) ? expf((row[c] - mx) * inv_t) : -1.f;
    key[c] = e;
    idx[c] = c;
    if (c < V) se += e;
  }
  se = warp_sum(se);
  if (lane == 0) red[w] = se;
  __syncthreads();
  se = 0.f;
#pragma unroll
  for (int i = 0; i < SMP_THREADS / 32; ++i) se += red[i];
  const float inv_se = 1.f / se;

  // ---- sort descending by key (ties: lower index first) by RANK COUNTING: element c goes to position
  //      #{j : key[j] > key[c] or (key[j] == key[c] and j < c)}.  V^2 / 256 broadcast smem reads per thread and
  //      two barriers instead of the ~50 barrier-separated passes of a bitonic network (the sampler sits on the
  //      critical path of every generated token) ----
  __syncthreads();
  {
    float myk[SMP_N / SMP_THREADS];
    int myr[SMP_N / SMP_THREADS];
#pragma unroll
    for (int i = 0; i < SMP_N / SMP_THREADS; ++i) {
      const int c = tid + i * SMP_THREADS;
      myk[i] = (c < V) ? key[c] : -1.f;
      myr[i] = 0;
    }
    for (int j = 0; j < V; ++j) {
      const float kj = key[j];
#pragma unroll
      for (int i = 0; i < SMP_N / SMP_THREADS; ++i) {
        const int c = tid + i * SMP_THREADS;
        myr[i] += (kj > myk[i] || (kj == myk[i] && j < c)) ? 1 : 0;
      }
    }
    __syncthreads();                       // everyone has read the unsorted keys
#pragma unroll
    for (int i = 0; i < SMP_N / SMP_THREADS; ++i) {
      const int c = tid + i * SMP_THREADS;
      if (c < V) { key[myr[i]] = myk[i]; idx[myr[i]] = c; }
    }
  }
  __syncthreads();

  // ---- inclusive scan of probabilities (4 consecutive per thread) ----
  float p4[4], run = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = tid * 4 + i;
    float pr = (c < V) ? key[c] * inv_se : 0.f;
    run += pr;
    p4[i] = run;
  }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  if (tid == 0) { s_first = SMP_N; s_choice = -1; }
  (void)s_choice;
  __syncthreads();
  float base = incl - run;
  for (int i = 0; i < w; ++i) base += wsum[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) { p4[i] += base; key[tid * 4 + i] = p4[i]; }   // key := cumulative mass
  // first sorted position whose cumulative mass exceeds top_p
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = tid * 4 + i;
    if (c < V && p4[i] > top_p) { atomicMin(&s_first, c); break; }
  }
  __syncthreads();
  int first = s_first, ncand, st = 0;
  if (first >= V) ncand = V < 3 ? V : 3;              // nothing above p: reference takes the top 3
  else if (first + 1 >= V) { ncand = V; st = 1; }     // exactly one index above: reference IndexError
  else ncand = first + 1;                             // `where(..)[0][1]` == first + 1 candidates
  float total = key[ncand - 1];
  if (banned) {
    // Grammar-constrained draw (opt-in): the reference draws from the nucleus candidates and REJECTS
    // inadmissible tokens (Beat going backwards, PAD, early EOS: inference.py:279-310), re-running the model and
    // re-drawing -- i.e. it samples from the candidate distribution restricted to the admissible tokens.  Here the
    // inadmissible candidates get zero mass and the cumulative sums are rebuilt over the SAME candidate set: the same
    // distribution in one draw.  status 2 = every candidate is inadmissible (the reference would spin to its 256-retry
    // abort).
    const uint8_t* brow = banned + (int64_t)r * V;
    __syncthreads();
    float run2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid * 4 + i;
      float prev = (c == 0) ? 0.f : key[c - 1];
      float pr = (c < ncand && !brow[idx[c]]) ? (p4[i] - prev) : 0.f;       // this candidate's own mass
      run2 += pr;
      p4[i] = run2;
    }
    __syncthreads();                                   // all reads of the unmasked cumulative sums are done
    float incl2 = run2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, incl2, o);
      if (lane >= o) incl2 += t;
    }
    if (lane == 31) wsum[w] = incl2;
    __syncthreads();
    float base2 = incl2 - run2;
    for (int i = 0; i < w; ++i) base2 += wsum[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { p4[i] += base2; key[tid * 4 + i] = p4[i]; }
    __syncthreads();
    total = key[ncand - 1];
    if (total <= 0.f) st = 2;
  }
  const float thresh = u[r] * total;         // cdf_i = cum_i / total > u  <=>  cum_i > u * total
  __syncthreads();
  if (tid == 0) s_first = ncand - 1;                  // fallback: last candidate
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = tid * 4 + i;
    if (c < ncand && p4[i] > thresh) { atomicMin(&s_first, c); break; }
  }
  __syncthreads();
  if (tid == 0) {
    out[r] = idx[s_first];
    if (status) status[r] = st;
  }
}


__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                             float inv_t, float top_p, const float* __restrict__ u,
                                                             int greedy, int64_t* __restrict__ out,
                                                             int32_t* __restrict__ status,
                                                             const uint8_t* __restrict__ banned,
                                                             const float* __restrict__ t_rows) {
  pdl_trigger();
  pdl_wait();                       // (no-ops unless launched as a dependent of the decode step, emo_set_pdl)
  sample_row(logits + (int64_t)blockIdx.x * ld, (int)blockIdx.x, V, inv_t, top_p, u, greedy, out, status, banned, t_rows);
}

// ---------------------------------------------------------------------------------------------
// Fused logits projection + sampler (north star: "a fused logits-projection + top-p / temperature sampler for the
// autoregressive decode loop"; stage2_accompaniment/inference.py:272-276, stage1_compose/inference_utils.py:66-72).
// One thread-block CLUSTER of LS_CL CTAs per sequence: every CTA streams its slice of the [V, 512] output weights
// (one warp per vocabulary entry, the arithmetic of gemm_skinny_nt_kernel instruction for instruction, so the logits --
// and with them the greedy tokens -- are bit-identical to the two-kernel path), optionally after the LayerNorm of the
// hidden row (the post-LN Performer's last norm2), writes its logits to HBM (the host-side reject-and-redraw loop reads
// them) AND into the shared memory of the cluster's CTA 0 through DSMEM; after one cluster barrier CTA 0 draws the
// token from its shared-memory copy.  The logits never make the HBM round trip between two kernels, and the step
// loses one dependent launch.
// ---------------------------------------------------------------------------------------------
constexpr int LS_CL = 8;
constexpr int LS_K = 512;
template <int NC>   // vocabulary rows per warp: all of them are requested before the kernel waits for its predecessor
__global__ void __launch_bounds__(SMP_THREADS) logits_sample_kernel(const bf16* __restrict__ x, int64_t ldx,
                                                                    const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                                    const bf16* __restrict__ W, int64_t ldw,
                                                                    const float* __restrict__ bias, int V,
                                                                    float* __restrict__ logits, int64_t ld, float inv_t,
                                                                    float top_p, const float* __restrict__ u, int greedy,
                                                                    int64_t* __restrict__ out, int32_t* __restrict__ status,
                                                                    const uint8_t* __restrict__ banned,
                                                                    const float* __restrict__ t_rows) {
  __shared__ __align__(16) bf16 xs[LS_K];
  __shared__ float lrow[SMP_N];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int r = blockIdx.y;
  const int per = (V + LS_CL - 1) / LS_CL;
  const int n_lo = (int)rank * per, n_hi = (n_lo + per < V) ? n_lo + per : V;
  constexpr int kv = LS_K / 8;
  // the weights do not depend on the step: every vocabulary row of this warp (NC x 1 KB) is in flight before the wait
  // on the previous kernel of the step, so that what remains on the token's critical path is the dot products
  uint4 bw[NC][2];
  float bias_n[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int n = n_lo + w + i * (SMP_THREADS / 32);
    bias_n[i] = 0.f;
    if (n < n_hi) {
      const uint4* brow = reinterpret_cast<const uint4*>(W + (int64_t)n * ldw);
      bw[i][0] = __ldg(brow + lane);
      bw[i][1] = __ldg(brow + lane + 32);
      if (bias) bias_n[i] = __ldg(bias + n);
    } else {
      bw[i][0] = make_uint4(0u, 0u, 0u, 0u);
      bw[i][1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float4 lg[4], lb[4];
  if (ln_g && w == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = (lane + 32 * h) * 8;
      lg[2 * h] = __ldg(reinterpret_cast<const float4*>(ln_g + c0));
      lg[2 * h + 1] = __ldg(reinterpret_cast<const float4*>(ln_g + c0 + 4));
      lb[2 * h] = __ldg(reinterpret_cast<const float4*>(ln_b + c0));
      lb[2 * h + 1] = __ldg(reinterpret_cast<const float4*>(ln_b + c0 + 4));
    }
  }
  pdl_trigger();
  pdl_wait();
  if (tid < kv) reinterpret_cast<uint4*>(xs)[tid] = *reinterpret_cast<const uint4*>(x + (int64_t)r * ldx + tid * 8);
  __syncthreads();
  if (ln_g) {
    if (w == 0) {       // the numbers emo_ln_fwd / the GEMV's LayerNorm prologue produce (fp32 statistics, bf16 result)
      uint4* row = reinterpret_cast<uint4*>(xs);
      float xv[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 t = row[lane + 32 * h];
        unpack_bf16x2(t.x, xv[8 * h], xv[8 * h + 1]); unpack_bf16x2(t.y, xv[8 * h + 2], xv[8 * h + 3]);
        unpack_bf16x2(t.z, xv[8 * h + 4], xv[8 * h + 5]); unpack_bf16x2(t.w, xv[8 * h + 6], xv[8 * h + 7]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) sum += xv[j];
      const float mu = warp_sum(sum) * (1.f / 512.f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { float d = xv[j] - mu; q += d * d; }
      const float rs = rsqrtf(warp_sum(q) * (1.f / 512.f) + 1e-5f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float y[8];
        const float gq[8] = {lg[2 * h].x, lg[2 * h].y, lg[2 * h].z, lg[2 * h].w, lg[2 * h + 1].x, lg[2 * h + 1].y, lg[2 * h + 1].z, lg[2 * h + 1].w};
        const float bq[8] = {lb[2 * h].x, lb[2 * h].y, lb[2 * h].z, lb[2 * h].w, lb[2 * h + 1].x, lb[2 * h + 1].y, lb[2 * h + 1].z, lb[2 * h + 1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (xv[8 * h + j] - mu) * rs * gq[j] + bq[j];
        uint4 t;
        t.x = pack_bf16x2(y[0], y[1]); t.y = pack_bf16x2(y[2], y[3]); t.z = pack_bf16x2(y[4], y[5]); t.w = pack_bf16x2(y[6], y[7]);
        row[lane + 32 * h] = t;
      }
    }
    __syncthreads();
  }
  // address of lrow[] in CTA 0 of the cluster
  uint32_t lrow0;
  {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(lrow);
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lrow0) : "r"(local), "r"(0));
  }
  float acc[NC];
  {
    uint4 aw[2];
    aw[0] = reinterpret_cast<const uint4*>(xs)[lane];
    aw[1] = reinterpret_cast<const uint4*>(xs)[lane + 32];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      acc[i] = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float b[8], a[8];
        unpack_bf16x2(bw[i][h].x, b[0], b[1]); unpack_bf16x2(bw[i][h].y, b[2], b[3]);
        unpack_bf16x2(bw[i][h].z, b[4], b[5]); unpack_bf16x2(bw[i][h].w, b[6], b[7]);
        unpack_bf16x2(aw[h].x, a[0], a[1]); unpack_bf16x2(aw[h].y, a[2], a[3]);
        unpack_bf16x2(aw[h].z, a[4], a[5]); unpack_bf16x2(aw[h].w, a[6], a[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i] = fmaf(a[j], b[j], acc[i]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {          // warp_sum of every row, interleaved
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int n = n_lo + w + i * (SMP_THREADS / 32);
      if (n < n_hi) {
        const float v = acc[i] + bias_n[i];
        logits[(int64_t)r * ld + n] = v;
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(lrow0 + 4u * (uint32_t)n), "f"(v) : "memory");
      }
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank != 0) return;
  sample_row(lrow, r, V, inv_t, top_p, u, greedy, out, status, banned, t_rows);
}

static int logits_sample_launch(const void* x, int64_t ldx, const float* ln_g, const float* ln_b, const void* W, int64_t ldw,
                                const float* bias, int rows, int V, float* logits, int64_t ld, float inv_t, const float* t_rows,
                                float top_p, const float* u, int greedy, int64_t* out, int32_t* status, const uint8_t* banned,
                                cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LS_CL, rows);
  cfg.blockDim = dim3(SMP_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = LS_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = emo_pdl_enabled() ? 2 : 1;
  // 8 CTAs x 8 warps: 6 vocabulary rows per warp cover V <= 384 (the three models: 329 / 372 / 216), 16 cover SMP_N
  if (V <= 6 * LS_CL * (SMP_THREADS / 32))
    EMO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, logits_sample_kernel<6>, (const bf16*)x, ldx, ln_g, ln_b, (const bf16*)W, ldw, bias, V, logits, ld,
                                      inv_t, top_p, u, greedy, out, status, banned, t_rows));
  else
    EMO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, logits_sample_kernel<16>, (const bf16*)x, ldx, ln_g, ln_b, (const bf16*)W, ldw, bias, V, logits, ld,
                                      inv_t, top_p, u, greedy, out, status, banned, t_rows));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

extern "C" int emo_logits_sample(const void* x, int64_t ldx, const float* ln_gamma, const float* ln_beta, const void* W,
                                 int64_t ldw, const float* bias, int rows, int V, int K, float* logits, int64_t ld,
                                 float temperature, const float* temperature_rows, float top_p, const float* u, int greedy,
                                 int64_t* out, int32_t* status, const uint8_t* banned, void* stream) {
  EMO_REQUIRE(V > 0 && V <= SMP_N, "emo_logits_sample: V must be in [1, %d]", SMP_N);
  EMO_REQUIRE(K == LS_K, "emo_logits_sample: K must be %d (d_model of the three models)", LS_K);
  EMO_REQUIRE(x && W && logits && out, "emo_logits_sample: null operand");
  EMO_REQUIRE((ldx % 8 == 0) && (ldw % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0),
              "emo_logits_sample: hidden rows / weights must be 16-byte aligned with leading dimensions that are multiples of 8");
  EMO_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "emo_logits_sample: ln_gamma and ln_beta come together");
  EMO_REQUIRE(greedy || (u != nullptr && (temperature_rows != nullptr || temperature > 0.f)),
              "emo_logits_sample: sampling needs u and a temperature > 0");
  if (rows == 0) return EMO_OK;
  const float inv_t = (greedy || temperature_rows) ? 1.f : 1.f / temperature;
  return logits_sample_launch(x, ldx, ln_gamma, ln_beta, W, ldw, bias, rows, V, logits, ld, inv_t, temperature_rows, top_p, u, greedy,
                              out, status, banned, (cudaStream_t)stream);
}

extern "C" int emo_sample(const float* logits, int64_t ld, int rows, int V, float temperature, float top_p,
                          const float* u, int greedy, int64_t* out, int32_t* status, const uint8_t* banned,
                          void* stream) {
  EMO_REQUIRE(V > 0 && V <= SMP_N, "emo_sample: V must be in [1, %d]", SMP_N);
  EMO_REQUIRE(greedy || (u != nullptr && temperature > 0.f), "emo_sample: sampling needs u and temperature > 0");
  if (rows == 0) return EMO_OK;
  EMO_CHECK_CUDA(emo_launch_dep(sample_kernel, dim3(rows), dim3(SMP_THREADS), 0, (cudaStream_t)stream, logits, ld, V,
                                greedy ? 1.f : 1.f / temperature, top_p, u, greedy, out, status, banned, (const float*)nullptr));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

extern "C" int emo_sample_rows(const float* logits, int64_t ld, int rows, int V, const float* temperature_rows, float top_p,
                               const float* u, int greedy, int64_t* out, int32_t* status, const uint8_t* banned,
                               void* stream) {
  EMO_REQUIRE(V > 0 && V <= SMP_N, "emo_sample_rows: V must be in [1, %d]", SMP_N);
  EMO_REQUIRE(greedy || (u != nullptr && temperature_rows != nullptr), "emo_sample_rows: sampling needs u and the temperatures");
  if (rows == 0) return EMO_OK;
  EMO_CHECK_CUDA(emo_launch_dep(sample_kernel, dim3(rows), dim3(SMP_THREADS), 0, (cudaStream_t)stream, logits, ld, V, 1.f, top_p, u,
                                greedy, out, status, banned, temperature_rows));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
