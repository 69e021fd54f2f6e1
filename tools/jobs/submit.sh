#!/bin/bash
# usage: tools/jobs/submit.sh <name> <timeout_s> <command...>   (retries while the pod answers busy; log -> gpurun_out/<name>_call.log)
name=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/${name}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${name}_call.log; then break; fi
  sleep 60
done
echo "submit rc=$rc" >> gpurun_out/${name}_call.log
