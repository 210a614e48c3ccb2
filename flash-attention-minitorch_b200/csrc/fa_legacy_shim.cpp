// fa_legacy_shim.cpp -- the reference's ctypes entry points, one tiny shared library per reference
// library name (reference loader: minitorch/cuda_kernel_ops.py:27-35).
//
// Compiled six times by compile_cuda.sh with -DFA_SHIM_CONVENTION={1,2} and -DFA_SHIM_FW / -DFA_SHIM_BW:
//   flash_attn_fw.so, flash_attn_causal_fw.so  -> launch_flash_attn_fw, FA-1 statistics (l = sum exp, m = max)
//   flash_attn2_fw.so                          -> launch_flash_attn_fw, FA-2 statistics (l = logsumexp)
//   flash_attn_bw.so, flash_attn_causal_bw.so  -> launch_flash_attn_bw consuming (l, m)
//   flash_attn2_bw.so                          -> launch_flash_attn_bw consuming L
// Signatures: reference src/flash_attn2_fw.cu:312-322, src/flash_attn2_bw.cu:279-293.  The "causal"
// pair differs from the plain FA-1 pair only by skipping masked tiles (-DCAUSAL_BLOCKSPARSE, reference
// Makefile:37-42), which the B200 kernels always do; results are identical.
// Error behaviour follows the reference: message on stderr, exit(EXIT_FAILURE).
#include <stdio.h>
#include <stdlib.h>

#include "../../include/fa_b200.h"

#ifndef FA_SHIM_CONVENTION
#error "define FA_SHIM_CONVENTION to 1 (FA-1) or 2 (FA-2)"
#endif

extern "C" {

#ifdef FA_SHIM_FW
void launch_flash_attn_fw(float* q, float* k, float* v, float* out, float* l, float* m, int batch, int N, int d,
                          bool causal_mask, void* stream) {
  int rc = fa_b200_host_fwd(q, k, v, out, l, m, batch, N, d, causal_mask ? 1 : 0, FA_SHIM_CONVENTION, stream);
  if (rc != FA_OK) {
    fprintf(stderr, "Flash Attention Error: %s\n", fa_b200_last_error());
    exit(EXIT_FAILURE);
  }
}
#endif

#ifdef FA_SHIM_BW
void launch_flash_attn_bw(float* q, float* k, float* v, float* out, float* out_grad, float* q_grad, float* k_grad,
                          float* v_grad, float* l, float* m, int batch, int N, int d, bool causal_mask,
                          void* stream) {
  int rc = fa_b200_host_bwd(q, k, v, out, out_grad, q_grad, k_grad, v_grad, l, m, batch, N, d, causal_mask ? 1 : 0,
                            FA_SHIM_CONVENTION, stream);
  if (rc != FA_OK) {
    fprintf(stderr, "Flash Attention Error: %s\n", fa_b200_last_error());
    exit(EXIT_FAILURE);
  }
}
#endif

}  // extern "C"
