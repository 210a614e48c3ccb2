#!/usr/bin/env bash
# tools/sustained_probe.sh <label> <command...>: runs the command while sampling SM clock and board power every 100 ms;
# prints the command's last line and the medians over the second half of the samples (the power-capped steady state).
label="$1"; shift
tmp=$(mktemp)
( while true; do nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits; sleep 0.1; done ) > "$tmp" &
spid=$!
out=$("$@" 2>&1 | tail -1)
kill $spid 2>/dev/null; wait $spid 2>/dev/null
python3 - "$tmp" "$label" "$out" <<'PY'
import sys, statistics
rows = [l.strip().split(',') for l in open(sys.argv[1]) if ',' in l]
rows = rows[len(rows) // 2:]
mhz = [float(r[0]) for r in rows]; w = [float(r[1]) for r in rows]
print(f"{sys.argv[2]:<26} {sys.argv[3]}   | steady state: {statistics.median(mhz):.0f} MHz, {statistics.median(w):.0f} W ({len(rows)} samples)")
PY
rm -f "$tmp"
