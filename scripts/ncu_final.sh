#!/bin/bash
# Final round captures at the bench's own size (1M objects / 1M rays): full sets for the dominant kernels.
TAG=${1:-r1f}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_cc_epa|k_cc_gjk|k_cc_manifold|k_pair_search|k_aabb|k_bh_epa|k_narrow<\(int\)5>" -s 21 -c 7 -f \
    -o gpurun_out/prof_${TAG}_world python bench.py --steps 1 --warmup 3 --no-cpu --no-rays > gpurun_out/ncu_${TAG}_world.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ray_cast" -s 3 -c 2 -f \
    -o gpurun_out/prof_${TAG}_rays python bench.py --rays-only --steps 3 --warmup 3 > gpurun_out/ncu_${TAG}_rays.log 2>&1
ls -la gpurun_out | tail -6
