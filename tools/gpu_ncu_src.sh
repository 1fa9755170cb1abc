#!/bin/bash
# source-level ncu capture (per-line instruction counts) of the hot kernels on the full C2 workload
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu ${BENCH_ARGS}"
$CMD > gpurun_out/plain_src.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_src.log; exit 1; }
for K in ${NCU_KERNELS:-parse_kernel radix_scatter_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:"^${K}" -s ${NCU_S:-1} -c 1 -f -o gpurun_out/src_${K} $CMD > gpurun_out/ncu_src_${K}.log 2>&1
  echo "ncu ${K} exit $?"; tail -2 gpurun_out/ncu_src_${K}.log
done
ls -la gpurun_out/*.ncu-rep
