#!/bin/bash
# usage: tools/profile_workload.sh <workload> <kernel-regex> [launch-skip] [launch-count]   (run on the GPU box, one per gpurun call)
# 1) the bench command exits 0 without ncu, 2) the same command under `ncu --set full` for the named kernels.
cd "$(dirname "$0")/.."
wl=$1; rx=$2; skip=${3:-6}; cnt=${4:-3}
python bench.py --workload $wl --steps 4 --warmup 3 --no-extras --no-cpu > gpurun_out/plain_$wl.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$wl.log; exit 1; }
tail -1 gpurun_out/plain_$wl.log | cut -c1-300
B200CA_BENCH_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/r02_$wl \
    python bench.py --workload $wl --steps 4 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_$wl.log 2>&1
tail -2 gpurun_out/ncu_$wl.log
