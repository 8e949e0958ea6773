#!/bin/bash
# Round 2, GPU session 14 (1 GPU): ncu full captures (with source) of the convex-convex kernels + pair search, and the launch list of the bench command.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_cc_gjk|k_cc_manifold|k_cc_epa_tier|k_pair_search" -s 12 -c 5 -f -o gpurun_out/r2n_prof_world \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --no-extras --no-traffic --no-secondary > gpurun_out/r2n_ncu.log 2>&1
ls -la gpurun_out/r2n_prof_world.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-traffic --no-secondary > gpurun_out/r2n_under_ncu.log 2>&1
wc -l gpurun_out/r2n_launches.csv
