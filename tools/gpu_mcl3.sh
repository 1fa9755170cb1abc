#!/bin/bash
# MCL GPU tests + timings after a kernel change
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_mcl_gpu.py -q -x -p no:cacheprovider > gpurun_out/mcl_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/mcl_gpu.log
timeout 120 python tools/mcl_timing.py 1.2 C2_small ${MCL_TIMING_EXTRA} > gpurun_out/mcl_timing.jsonl 2> gpurun_out/mcl_timing.err; echo "timing exit $?"; cat gpurun_out/mcl_timing.jsonl; tail -3 gpurun_out/mcl_timing.err
