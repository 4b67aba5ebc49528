// K2 (bf16 tensor-core mode) -- placeholder until the tcgen05 kernels land.
#include "fecl_internal.h"

namespace dycon {
size_t fecl_tc_state_bytes(int, int, int, int) { return 0; }
size_t fecl_tc_workspace_bytes(int, int, int) { return 0; }
int fecl_tc_fwd(const FeclProblem&, const FeclFwdArgs&, cudaStream_t) {
  return fail(DYCON_ERR_UNSUPPORTED, "FeCL bf16 (tcgen05) path not built yet");
}
int fecl_tc_bwd(const FeclProblem&, const FeclBwdArgs&, cudaStream_t) {
  return fail(DYCON_ERR_UNSUPPORTED, "FeCL bf16 (tcgen05) path not built yet");
}
}  // namespace dycon
