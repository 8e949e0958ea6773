#!/bin/bash
# Round 2, GPU session 7 (1 GPU): routed sharding replayed on one device.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_routed_shards.py -x -q -m gpu > gpurun_out/r2g_routed.log 2>&1; echo "routed exit $?" >> gpurun_out/r2g_routed.log
tail -40 gpurun_out/r2g_routed.log
