// fa_bwd_sm100.cu -- flash-attention backward for sm_100a (tcgen05 + TMEM + TMA).  [placeholder: round-1 WIP]
#include "fa_internal.h"

namespace fa {

bool tc_bwd_supported(int) { return false; }
size_t tc_bwd_workspace_bytes(int, int, int) { return 0; }
cudaError_t tc_bwd(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*,
                   const __nv_bfloat16*, const float*, void*, void*, void*, bool, bool, void*, int, int, int, bool,
                   cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace fa
