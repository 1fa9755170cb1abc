#!/bin/bash
# MCL on the device: GPU parity tests, the smoke run, then timings on device-resident graphs
mkdir -p gpurun_out
timeout ${MCL_TEST_TIMEOUT:-200} python -m pytest tests/test_zz_mcl_gpu.py -q -rA -s --timeout 120 -p no:cacheprovider > gpurun_out/mcl_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|PASSED|FAILED|bit-equal|Error|assert " gpurun_out/mcl_gpu.log | tail -40
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout ${MCL_TIMING_TIMEOUT:-120} python tools/mcl_timing.py 1.2 C2_small ${MCL_TIMING_EXTRA} > gpurun_out/mcl_timing.jsonl 2> gpurun_out/mcl_timing.err; echo "timing exit $?"; cat gpurun_out/mcl_timing.jsonl; tail -3 gpurun_out/mcl_timing.err
