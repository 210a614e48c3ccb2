#!/usr/bin/env bash
# tools/gpurun_retry.sh <log> <gpurun args...>: retry while the pod answers "transient" (exit code 3), up to 20 times
log="$1"; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 60
done
exit 3
