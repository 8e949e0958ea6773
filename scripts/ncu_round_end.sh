#!/bin/bash
# Round-end ncu evidence at the bench's own size (run under gpurun from the repo root): bash scripts/ncu_round_end.sh <tag>
#  (1) launch list of one bench step (cold-cache, serialised: compare SHARES), (2) full sets for the dominant kernels.
TAG=${1:-r1end}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cc_epa|k_cc_gjk|k_cc_manifold|k_pair_search|k_aabb|k_bh_epa" -s 18 -c 6 -f \
    -o gpurun_out/prof_${TAG}_world python bench.py --steps 1 --warmup 3 --no-cpu --no-rays > gpurun_out/ncu_${TAG}_world.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ray_cast" -s 3 -c 2 -f \
    -o gpurun_out/prof_${TAG}_rays python bench.py --rays-only --steps 3 --warmup 3 > gpurun_out/ncu_${TAG}_rays.log 2>&1
ls -la gpurun_out | tail -6
