#!/usr/bin/env bash
# compile_cuda.sh -- build libfa_b200.so (sm_100a) and the six reference-named shim libraries.
#
# Counterpart of the reference's compile_cuda.sh / Makefile:14-50.  Output layout:
#   flash-attention-minitorch_b200/lib/libfa_b200.so
#   flash-attention-minitorch_b200/minitorch/cuda_kernels/{flash_attn_fw,flash_attn_bw,flash_attn_causal_fw,
#       flash_attn_causal_bw,flash_attn2_fw,flash_attn2_bw}.so      (what minitorch/cuda_kernel_ops.py:30-35 loads)
# Usage: ./compile_cuda.sh [-j N]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/csrc"
OBJ="$HERE/build"
LIB="$HERE/lib"
SHIM="$HERE/minitorch/cuda_kernels"
NVCC="${NVCC:-nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v ${FA_EXTRA_NVFLAGS:-}"
mkdir -p "$OBJ" "$LIB" "$SHIM"

pids=()
for f in fa_simt fa_util fa_capi fa_fwd_sm100 fa_fwd_pair_sm100 fa_bwd_sm100 fa_bwd_tf32_sm100 fa_gemm_sm100 fa_lm_kernels; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$SRC/$f.cu" -nt "$OBJ/$f.o" ] || [ -n "$(find "$SRC" "$HERE/../include" -name '*.h' -newer "$OBJ/$f.o" -o -name '*.cuh' -newer "$OBJ/$f.o")" ]; then
    ( $NVCC $NVFLAGS -c "$SRC/$f.cu" -o "$OBJ/$f.o" > "$OBJ/$f.ptxas.log" 2>&1 || { cat "$OBJ/$f.ptxas.log"; exit 1; } ) &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do [ -n "$p" ] && { wait "$p" || fail=1; }; done
[ "$fail" = 0 ] || { echo "compile_cuda.sh: compilation failed" >&2; exit 1; }

g++ -O2 -std=c++17 -fPIC -I/usr/local/cuda/include -c "$SRC/fa_nccl.cpp" -o "$OBJ/fa_nccl.o"
$NVCC $ARCH -shared -o "$LIB/libfa_b200.so" "$OBJ"/fa_nccl.o "$OBJ"/fa_simt.o "$OBJ"/fa_util.o "$OBJ"/fa_capi.o "$OBJ"/fa_fwd_sm100.o "$OBJ"/fa_fwd_pair_sm100.o \
  "$OBJ"/fa_bwd_sm100.o "$OBJ"/fa_bwd_tf32_sm100.o "$OBJ"/fa_gemm_sm100.o "$OBJ"/fa_lm_kernels.o -lcudart -ldl

shim() {  # name convention FW|BW
  g++ -O2 -fPIC -shared -DFA_SHIM_CONVENTION="$2" -DFA_SHIM_"$3" "$SRC/fa_legacy_shim.cpp" \
    -o "$SHIM/$1.so" -L"$LIB" -lfa_b200 -Wl,-rpath,'$ORIGIN/../../lib'
}
shim flash_attn_fw 1 FW
shim flash_attn_bw 1 BW
shim flash_attn_causal_fw 1 FW
shim flash_attn_causal_bw 1 BW
shim flash_attn2_fw 2 FW
shim flash_attn2_bw 2 BW
echo "built $LIB/libfa_b200.so and 6 shims in $SHIM"
