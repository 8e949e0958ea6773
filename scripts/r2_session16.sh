#!/bin/bash
# 2-D Polyline ray casting on the GPU: parity tests (2-D and the refactored 3-D traversal) and the ray figures.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rays2d.py tests/test_rays_r2.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --rays-only --no-cpu > gpurun_out/rays_only.json 2> gpurun_out/rays_only.err
tail -c 3000 gpurun_out/rays_only.json
tail -5 gpurun_out/rays_only.err
