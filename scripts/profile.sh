#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) full capture of the top kernels.
# Run under gpurun from the repo root: bash scripts/profile.sh <tag> [n_objects_for_full_capture]
TAG=${1:-r1}
NFULL=${2:-300000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_narrow|k_gjk|k_epa|k_manifold|k_pair_search|k_ray_cast" -s 40 -c 16 -f \
    -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu --n-objects ${NFULL} > gpurun_out/bench_under_ncu2_${TAG}.log 2>&1
ls -la gpurun_out
