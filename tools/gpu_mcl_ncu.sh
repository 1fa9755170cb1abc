#!/bin/bash
# ncu evidence for the MCL kernels on a small graph (C2_tiny, 2 400 genes): launch list of a whole run, then a full capture of the two
# expansion kernels and the inflation kernel.  Run only after tools/gpu_mcl.sh exited 0 without ncu.
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/mcl_launches_C2_tiny.csv python tools/mcl_timing.py 1.2 C2_tiny > gpurun_out/ncu_mcl_list.log 2>&1; echo "ncu list exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"mcl_expand_direct_kernel|mcl_expand_kernel|mcl_inflate_kernel" -s ${NCU_S:-12} -c ${NCU_C:-6} -f -o gpurun_out/mcl_prof python tools/mcl_timing.py 1.2 C2_tiny > gpurun_out/ncu_mcl_full.log 2>&1; echo "ncu full exit $?"; ls -la gpurun_out | tail -6
