// K8, decode step: one new token per sequence of a RAGGED batch against each sequence's own K|V cache
// (stage2_accompaniment/inference.py:252-272 re-runs HF GPT2Attention over the whole prefix for every token; here the
// prefix lives in the cache).  One static-shaped launch for the whole batch -- the per-sequence lengths come from a
// device array -- so the GPT-2 step can be captured in a CUDA graph like the Performer step.
// grid (H, B), 128 threads: append this head's k | v row at pos[b], scores of the new query against keys [0, pos[b]],
// exact fp32 softmax, weighted sum of the values.  Latency-bound by design (a few thousand MACs per block).
#include "common.cuh"

namespace {
constexpr int DE = 64;

template <typename T>
__global__ void __launch_bounds__(128) attn_decode_kernel(const T* __restrict__ qkv, int64_t ld_qkv, T* __restrict__ kv, int64_t max_len,
                                                          const int64_t* __restrict__ pos, T* __restrict__ out, int64_t ld_out, int H,
                                                          float scale) {
  extern __shared__ float sc[];                    // [Tk] scores, then probabilities
  __shared__ float qs[DE], red[4], part[4][DE];
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, d = H * DE;
  const int64_t p = pos[b];
  const int Tk = (int)p + 1;
  const T* row = qkv + (int64_t)b * ld_qkv + h * DE;
  T* cache = kv + (int64_t)b * max_len * 2 * d;
  if (tid < DE) {
    qs[tid] = to_f(row[tid]) * scale;
    cache[p * 2 * d + h * DE + tid] = row[d + tid];            // K row of this head
  } else {
    cache[p * 2 * d + d + h * DE + (tid - DE)] = row[2 * d + (tid - DE)];   // V row
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int t = tid; t < Tk; t += 128) {
    const T* kr = cache + (int64_t)t * 2 * d + h * DE;
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < DE; e += Vec<T>::N) {
      Vec<T> v;
      v.load(kr + e);
#pragma unroll
      for (int j = 0; j < Vec<T>::N; ++j) acc += v.v[j] * qs[e + j];
    }
    sc[t] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int t = tid; t < Tk; t += 128) {
    const float e = expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  // values: lane c owns head columns 2c, 2c + 1; warp g takes keys g, g + 4, ...
  const int c = tid & 31, g = tid >> 5;
  float a0 = 0.f, a1 = 0.f;
  for (int t = g; t < Tk; t += 4) {
    const T* vr = cache + (int64_t)t * 2 * d + d + h * DE + 2 * c;
    const float pt = sc[t];
    a0 += pt * to_f(vr[0]);
    a1 += pt * to_f(vr[1]);
  }
  part[g][2 * c] = a0;
  part[g][2 * c + 1] = a1;
  __syncthreads();
  pdl_trigger();
  if (tid < DE) {
    const float o = (part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid]) * inv;
    out[(int64_t)b * ld_out + h * DE + tid] = from_f<T>(o);
  }
}
// Stage-1 (Transformer-XL) decode step: the same for the relative-position attention of optimus_txl_decoder.py:305-387.
// The reference keeps the last mem_len layer INPUTS as memory and re-derives K / V of every memory row on every step;
// K / V of a position never change (no dropout at inference), so they are cached instead and the new token attends
// over the last mem_len + 1 positions.  score_j = ((q + r_w_bias) . k_j + (q + r_r_bias) . r[p - j]) * scale, with
// r[dist] = r_net(pos_emb(dist)) precomputed per layer (rtab [mem_len + 1][H*64], indexed by DISTANCE).
template <typename T>
__global__ void __launch_bounds__(128) relattn_decode_kernel(const T* __restrict__ qkv, int64_t ld_qkv, T* __restrict__ kv, int64_t cap,
                                                             const int64_t* __restrict__ pos, const T* __restrict__ rtab,
                                                             const float* __restrict__ rwb, const float* __restrict__ rrb, int mem_len,
                                                             T* __restrict__ out, int64_t ld_out, int H, float scale) {
  extern __shared__ float sc[];                    // [mem_len + 1] scores, then probabilities
  __shared__ float qw[DE], qr[DE], red[4], part[4][DE];
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, d = H * DE;
  const int64_t p = pos[b];
  const int64_t j_lo = p > mem_len ? p - mem_len : 0;
  const int n = (int)(p - j_lo) + 1;
  const T* row = qkv + (int64_t)b * ld_qkv + h * DE;
  T* cache = kv + (int64_t)b * cap * 2 * d;
  if (tid < DE) {
    const float qv = to_f(row[tid]);
    qw[tid] = (qv + rwb[h * DE + tid]) * scale;
    qr[tid] = (qv + rrb[h * DE + tid]) * scale;
    cache[p * 2 * d + h * DE + tid] = row[d + tid];
  } else {
    cache[p * 2 * d + d + h * DE + (tid - DE)] = row[2 * d + (tid - DE)];
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int t = tid; t < n; t += 128) {
    const int64_t j = j_lo + t;
    const T* kr = cache + j * 2 * d + h * DE;
    const T* rr_ = rtab + (p - j) * d + h * DE;
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < DE; e += Vec<T>::N) {
      Vec<T> kv_, rv;
      kv_.load(kr + e);
      rv.load(rr_ + e);
#pragma unroll
      for (int i = 0; i < Vec<T>::N; ++i) acc += kv_.v[i] * qw[e + i] + rv.v[i] * qr[e + i];
    }
    sc[t] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int t = tid; t < n; t += 128) {
    const float e = expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  // the reference renormalises the (undropped) probabilities by (sum + 1e-8): sum = 1 here
  const float inv = 1.f / ((red[0] + red[1] + red[2] + red[3]) * (1.f + 1e-8f));
  const int c = tid & 31, g = tid >> 5;
  float a0 = 0.f, a1 = 0.f;
  for (int t = g; t < n; t += 4) {
    const T* vr = cache + (j_lo + t) * 2 * d + d + h * DE + 2 * c;
    const float pt = sc[t];
    a0 += pt * to_f(vr[0]);
    a1 += pt * to_f(vr[1]);
  }
  part[g][2 * c] = a0;
  part[g][2 * c + 1] = a1;
  __syncthreads();
  pdl_trigger();
  if (tid < DE) out[(int64_t)b * ld_out + h * DE + tid] = from_f<T>((part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid]) * inv);
}
}  // namespace

extern "C" int emo_relattn_decode_step(const void* qkv, int64_t ld_qkv, void* kv_cache, int64_t cap, const int64_t* pos, const void* rtab,
                                       const float* r_w_bias, const float* r_r_bias, int mem_len, void* out, int64_t ld_out, int B, int H,
                                       float scale, int dtype, void* stream) {
  EMO_REQUIRE(dtype == EMO_BF16 || dtype == EMO_F32, "emo_relattn_decode_step: bad dtype %d", dtype);
  const int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE((((uintptr_t)kv_cache) & 15) == 0 && (((uintptr_t)rtab) & 15) == 0 && (((int64_t)H * 64 * esz) % 16) == 0,
              "emo_relattn_decode_step: cache / table must be 16-byte aligned");
  EMO_REQUIRE(mem_len >= 0 && mem_len <= 24000 && cap > 0 && r_w_bias && r_r_bias && rtab, "emo_relattn_decode_step: bad arguments");
  if (B * H == 0) return EMO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)(mem_len + 1) * sizeof(float);
  static size_t configured[2] = {0, 0};
  if (smem > 40 * 1024 && smem > configured[dtype == EMO_BF16]) {
    if (dtype == EMO_BF16) EMO_CHECK_CUDA(cudaFuncSetAttribute(relattn_decode_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else EMO_CHECK_CUDA(cudaFuncSetAttribute(relattn_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dtype == EMO_BF16] = smem;
  }
  dim3 grid(H, B);
  if (dtype == EMO_BF16)
    EMO_CHECK_CUDA(emo_launch_dep(relattn_decode_kernel<bf16>, grid, dim3(128), smem, s, (const bf16*)qkv, ld_qkv, (bf16*)kv_cache, cap, pos,
                                  (const bf16*)rtab, r_w_bias, r_r_bias, mem_len, (bf16*)out, ld_out, H, scale));
  else
    EMO_CHECK_CUDA(emo_launch_dep(relattn_decode_kernel<float>, grid, dim3(128), smem, s, (const float*)qkv, ld_qkv, (float*)kv_cache, cap, pos,
                                  (const float*)rtab, r_w_bias, r_r_bias, mem_len, (float*)out, ld_out, H, scale));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}

extern "C" int emo_attn_decode_step(const void* qkv, int64_t ld_qkv, void* kv_cache, int64_t max_len, const int64_t* pos, void* out,
                                    int64_t ld_out, int B, int H, float scale, int dtype, void* stream) {
  EMO_REQUIRE(dtype == EMO_BF16 || dtype == EMO_F32, "emo_attn_decode_step: bad dtype %d", dtype);
  const int esz = dtype == EMO_BF16 ? 2 : 4;
  EMO_REQUIRE((((uintptr_t)kv_cache) & 15) == 0 && (((int64_t)H * 64 * esz) % 16) == 0, "emo_attn_decode_step: cache must be 16-byte aligned");
  EMO_REQUIRE(max_len > 0 && max_len <= 24000, "emo_attn_decode_step: max_len out of range");
  if (B * H == 0) return EMO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)max_len * sizeof(float);
  static size_t configured[2] = {0, 0};
  if (smem > 48 * 1024 && smem > configured[dtype == EMO_BF16]) {
    if (dtype == EMO_BF16) EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else EMO_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dtype == EMO_BF16] = smem;
  }
  dim3 grid(H, B);
  if (dtype == EMO_BF16)
    EMO_CHECK_CUDA(emo_launch_dep(attn_decode_kernel<bf16>, grid, dim3(128), smem, s, (const bf16*)qkv, ld_qkv, (bf16*)kv_cache, max_len, pos,
                                  (bf16*)out, ld_out, H, scale));
  else
    EMO_CHECK_CUDA(emo_launch_dep(attn_decode_kernel<float>, grid, dim3(128), smem, s, (const float*)qkv, ld_qkv, (float*)kv_cache, max_len, pos,
                                  (float*)out, ld_out, H, scale));
  EMO_LAUNCH_CHECK();
  return EMO_OK;
}
