#!/bin/bash
# round-end evidence: tests, bench (both arms), ncu launch list and full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
(nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2) > gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 1500 python bench.py --steps 5 --warmup 3 --e2e-check > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 exit $?"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"parse_kernel|radix_scatter_kernel|fit_kernel|rowsort_kernel|bin_cut_kernel" -s 5 -c 7 -f -o gpurun_out/prof_top_c2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | head -30
