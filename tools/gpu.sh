#!/bin/bash
# usage: tools/gpu.sh <tag> <timeout_s> <command...>   - retries while the pod answers busy (exit 3); log in gpurun_out/<tag>_call.log
tag=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${tag}_call.log; then break; fi
  sleep 60
done
echo "gpurun rc=$rc tries=$i" >> gpurun_out/${tag}_call.log
