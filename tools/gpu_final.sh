#!/bin/bash
# round-end evidence: ncu launch list and full capture of the top kernels on the C2 workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
(nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2) > gpurun_out/host.txt
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"^(parse_kernel|bk_binhist_kernel|rowsort_kernel|fit_kernel|scatter_rows_kernel|bin_cut_kernel|bk_cand_kernel|emit_kernel|scale_bh_kernel|connect_kernel)" -s 11 -c 11 -f -o gpurun_out/prof_top_c2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | grep -E "prof_top|launches"
