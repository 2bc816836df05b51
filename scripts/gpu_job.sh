#!/bin/bash
# Freezes the working tree under .frozen/<tag>/ and runs <command> from there on a GPU box, retrying while
# the pool is busy.  Lets development continue in the live tree while a job waits in the queue.
# usage: scripts/gpu_job.sh <tag> <timeout_s> [--gpus N] -- '<command>'
set -u
tag=$1; tmo=$2; shift 2
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
cmd=$1
root=/root/repo
rm -rf $root/.frozen/$tag
mkdir -p $root/.frozen/$tag
(cd $root && tar --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude='*.o' --exclude=__pycache__ --exclude=.pytest_cache -cf - .) | tar -xf - -C $root/.frozen/$tag
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $tmo "${extra[@]}" -- "cd .frozen/$tag && mkdir -p gpurun_out && ( $cmd ); rc=\$?; mkdir -p \$GRAFT_REPO_ROOT/gpurun_out/$tag; cp -r gpurun_out/. \$GRAFT_REPO_ROOT/gpurun_out/$tag/; exit \$rc" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient"; then
    echo "[gpu_job] attempt $attempt: pool busy, retrying in 120 s"; sleep 120; continue
  fi
  break
done
rm -rf $root/.frozen/$tag
