#!/usr/bin/env bash
# tools/build_variant.sh <name> [nvcc -D flags...]: builds the library with extra flags into build/variants/<name>/ for tools/ab.sh
set -euo pipefail
name="$1"; shift
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
tmp="$root/build/vtree_$name"
rm -rf "$tmp"; mkdir -p "$tmp/flash-attention-minitorch_b200" "$root/build/variants/$name"
cp -r "$root/flash-attention-minitorch_b200/csrc" "$root/flash-attention-minitorch_b200/compile_cuda.sh" "$tmp/flash-attention-minitorch_b200/"
cp -r "$root/include" "$tmp/"
(cd "$tmp/flash-attention-minitorch_b200" && FA_EXTRA_NVFLAGS="$*" bash compile_cuda.sh > "$root/build/variants/$name/build.log" 2>&1) || { tail -20 "$root/build/variants/$name/build.log"; exit 1; }
cp "$tmp/flash-attention-minitorch_b200/lib/libfa_b200.so" "$root/build/variants/$name/"
grep -A2 "fa_bwd_kernelILi128E13__nv_bfloat16Lb0" "$tmp/flash-attention-minitorch_b200/build/fa_bwd_sm100.ptxas.log" | grep spill | sed "s/^/$name: /"
rm -rf "$tmp"
