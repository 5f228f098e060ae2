// tcgen05 fused MLP -- placeholder until the kernel lands (algo 2 reports unsupported).
#include "common.cuh"
namespace cfm {
struct MlpBlobHeader;
size_t mlp_tc_blob_bytes(int, int, int) { return 0; }
int mlp_tc_prepare(const MlpBlobHeader&, void*, cudaStream_t) { return CFM_OK; }
int mlp_tc_supported(int, int, int, int) { return 0; }
size_t mlp_tc_workspace_bytes(int, int, int, int) { return 0; }
int mlp_tc_forward(const MlpBlobHeader&, const void*, const float*, int, const float*, float, int, float*,
                   void*, size_t, cudaStream_t) {
  set_error("mlp tcgen05 path not built");
  return CFM_ERR_ARG;
}
}  // namespace cfm
