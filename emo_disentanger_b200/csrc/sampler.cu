// K11: temperature + nucleus (top-p) sampling on device, one CTA per row (V <= 1024).
// Mirrors stage2_accompaniment/inference.py:71-100 / stage1_compose/inference_utils.py:14-41:
// softmax(l/t) -> sort descending -> cumulative sum -> cut at the SECOND index above top_p (the
// first crossing token is kept) -> renormalise -> inverse-CDF draw with the caller's uniform.
// greedy mode = argmax with numpy's lowest-index tie break (bit-exact decode mode).
#include "common.cuh"

constexpr int SMP_N = 1024;
constexpr int SMP_THREADS = 256;

__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                             float inv_t, float top_p, const float* __restrict__ u,
                                                             int greedy, int64_t* __restrict__ out,
                                                             int32_t* __restrict__ status,
                                                             const uint8_t* __restrict__ banned,
                                                             const float* __restrict__ t_rows) {
  __shared__ float key[SMP_N];
  __shared__ int idx[SMP_N];
  __shared__ float red[SMP_THREADS / 32];
  __shared__ int redi[SMP_THREADS / 32];
  __shared__ float wsum[SMP_THREADS / 32];
  __shared__ int s_first, s_choice;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* row = logits + (int64_t)blockIdx.x * ld;
  pdl_trigger();
  pdl_wait();                       // (no-ops unless launched as a dependent of the decode step, emo_set_pdl)
  if (t_rows && !greedy) inv_t = 1.f / t_rows[blockIdx.x];      // one temperature per row (a batch of quadrants)

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
    if (tid == 0) { out[blockIdx.x] = am; if (status) status[blockIdx.x] = 0; }
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
    const uint8_t* brow = banned + (int64_t)blockIdx.x * V;
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
  const float thresh = u[blockIdx.x] * total;         // cdf_i = cum_i / total > u  <=>  cum_i > u * total
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
    out[blockIdx.x] = idx[s_first];
    if (status) status[blockIdx.x] = st;
  }
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
