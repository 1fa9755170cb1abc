#!/bin/bash
# MCL on the device on the GPU box: parity tests, smoke, whole-process timing on a device-resident graph, a bench line with
# the opt-in MCL block.   gpurun --timeout 300 -- 'PYTHONUNBUFFERED=1 bash tools/gpu_mcl.sh'
mkdir -p gpurun_out
timeout ${MCL_TEST_TIMEOUT:-200} python -m pytest tests/test_zz_mcl_gpu.py -q -rA -s --timeout 120 -p no:cacheprovider > gpurun_out/mcl_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|bit-equal|Error|assert " gpurun_out/mcl_gpu.log | tail -30
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout ${MCL_TIMING_TIMEOUT:-120} python tools/mcl_timing.py 1.2 C2_small ${MCL_TIMING_EXTRA} > gpurun_out/mcl_timing.jsonl 2> gpurun_out/mcl_timing.err; echo "timing exit $?"; cat gpurun_out/mcl_timing.jsonl; tail -3 gpurun_out/mcl_timing.err
if [ -n "${MCL_BENCH}" ]; then
  timeout 120 python bench.py --config C2_small --steps 3 --warmup 3 --mcl 1.2 > gpurun_out/bench_C2_small_mcl.json 2> gpurun_out/bench_C2_small_mcl.err; echo "bench exit $?"
fi
