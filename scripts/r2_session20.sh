#!/bin/bash
# Per-kernel durations of one step (ncu, serialised) and a full capture of k_cc_manifold after the scan change.
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_cc_|k_pair_search|k_narrow|k_bh_epa" -s 40 -c 40 --csv --log-file gpurun_out/s20_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --no-rays --no-secondary --no-traffic > gpurun_out/s20_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cc_manifold|k_cc_epa_tier" -s 6 -c 3 -f -o gpurun_out/s20_prof_man python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-rays --no-secondary --no-traffic > gpurun_out/s20_b.log 2>&1
tail -2 gpurun_out/s20_b.log
