#!/bin/bash
# whole GPU suite, MCL timing on a device-resident graph, ncu launch list and one full capture of the MCL kernels (small graph)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_all.log
timeout 150 python tools/mcl_timing.py 1.2 C2_small > gpurun_out/mcl_timing.jsonl 2> gpurun_out/mcl_timing.err; echo "timing exit $?"; cat gpurun_out/mcl_timing.jsonl; tail -3 gpurun_out/mcl_timing.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/mcl_launches_C2_tiny.csv python tools/mcl_timing.py 1.2 C2_tiny > gpurun_out/ncu_mcl_list.log 2>&1; echo "ncu list exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"mcl_expand_kernel|mcl_inflate_kernel" -s 6 -c 4 -f -o gpurun_out/mcl_prof python tools/mcl_timing.py 1.2 C2_tiny > gpurun_out/ncu_mcl_full.log 2>&1; echo "ncu full exit $?"; ls -la gpurun_out | tail -8
