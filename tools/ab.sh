#!/usr/bin/env bash
# Same-box A/B of library builds: tools/ab.sh "<fa_check args>" variantA variantB ...   (build/variants/<name>/libfa_b200.so)
# Rounds are interleaved so that clock / thermal drift hits every variant alike.
args="$1"; shift
for round in 1 2 3; do
  for v in "$@"; do
    printf "%-10s " "$v"; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check $args 2>&1 | tail -1
  done
done
