// tcgen05 / TMEM / TMA bf16 MLP path (placeholder until the fused kernels land in this file).
#include "mlp_tc.cuh"

namespace knerf {

bool tc_path_compiled() { return false; }
int64_t tc_workspace_bytes(const Model&, int64_t, bool) { return -1; }
int64_t tc_packed_weight_bytes(const Model&) { return -1; }
int tc_pack_weights(const Model&, const float*, void*, cudaStream_t) {
  return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 path not built");
}
int tc_forward(const Model&, const float*, const void*, const float*, const float*, const float*, int64_t, int, bool,
               float*, char*, int64_t, cudaStream_t) {
  return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 path not built");
}
int tc_backward(const Model&, const float*, const void*, const float*, int64_t, int, float*, char*, int64_t,
                cudaStream_t) {
  return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 path not built");
}

}  // namespace knerf
