#!/bin/bash
# submit a gpurun call, retrying while the pod answers "busy" (exit 3: nothing charged)
# usage: tools/gpu/submit.sh [gpurun options] -- <command>
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_submit.log 2>&1; rc=$?
  if grep -q "status=transient" /tmp/gpurun_submit.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -12 /tmp/gpurun_submit.log
exit $rc
