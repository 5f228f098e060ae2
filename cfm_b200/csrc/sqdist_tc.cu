// tcgen05 3xTF32 cost-matrix GEMM -- placeholder until the kernel lands (algo 2 reports unsupported).
#include "common.cuh"
namespace cfm {
int sqdist_tc_supported(int, int, int, const float*, const float*, const float*, int64_t) { return 0; }
size_t sqdist_tc_workspace_bytes(int, int, int) { return 0; }
int sqdist_tc_launch(const float*, const float*, float*, int, int, int, int64_t, int, float*, const float*,
                     const float*, void*, size_t, cudaStream_t) {
  set_error("sqdist tcgen05 path not built");
  return CFM_ERR_ARG;
}
}  // namespace cfm
