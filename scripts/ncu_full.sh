#!/bin/bash
# Full ncu capture of selected kernels: bash scripts/ncu_full.sh <tag> <kernel-regex> <skip> <count> [n_objects]
TAG=$1; RE=$2; SKIP=${3:-0}; CNT=${4:-6}; N=${5:-300000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --n-objects $N > gpurun_out/ncu_${TAG}.log 2>&1
ls -la gpurun_out | tail -5
