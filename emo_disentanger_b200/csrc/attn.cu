// K8/K9 placeholders (implemented next): causal softmax attention and rel-pos attention.
#include "common.cuh"
extern "C" int emo_attn_fwd(const void*, const void*, const void*, int64_t, int64_t, void*, int64_t, float*, int, int, int, int, float, float, uint64_t, int, void*) {
  emo_set_error("emo_attn_fwd: not implemented yet"); return EMO_ERR_UNSUPPORTED; }
extern "C" int emo_attn_bwd(const void*, const void*, const void*, int64_t, int64_t, const void*, const void*, int64_t, const float*, void*, void*, void*, int64_t, int64_t, int, int, int, int, float, float, uint64_t, int, void*) {
  emo_set_error("emo_attn_bwd: not implemented yet"); return EMO_ERR_UNSUPPORTED; }
extern "C" int emo_relattn_fwd(const void*, const void*, const void*, int64_t, int64_t, const void*, int64_t, const float*, const float*, void*, int64_t, float*, int, int, int, int, float, int, void*) {
  emo_set_error("emo_relattn_fwd: not implemented yet"); return EMO_ERR_UNSUPPORTED; }
extern "C" int emo_relattn_bwd(const void*, const void*, const void*, int64_t, int64_t, const void*, int64_t, const float*, const float*, const void*, const void*, int64_t, const float*, void*, void*, void*, int64_t, int64_t, float*, float*, float*, int, int, int, int, float, int, void*) {
  emo_set_error("emo_relattn_bwd: not implemented yet"); return EMO_ERR_UNSUPPORTED; }
