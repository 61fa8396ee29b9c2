// Gram matrix on tensor cores (placeholder entry until the MN-major tcgen05 kernel lands).
#include "tc_common.cuh"
namespace fnst {
int gram_tc(const void* feat, float* out, int n, int hw, int c, int dtype, int device, cudaStream_t st) {
  set_error("gram: tensor-core path not built in this version");
  return -2;
}
}  // namespace fnst
