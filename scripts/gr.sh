#!/bin/bash
# usage: scripts/gr.sh <tag> <timeout_s> '<command>'  - runs gpurun from the repo root, retrying while the pod has no free slot
cd "$(dirname "$0")/.." || exit 1
tag=$1; tmo=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun --timeout "$tmo" "$@" > "/tmp/gr_$tag.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 60
done
echo "rc=$rc" >> "/tmp/gr_$tag.log"
