#!/bin/bash
# ncu capture of the top kernels on a small config (one ncu run per call)
mkdir -p gpurun_out
CMD="python bench.py --config C2_small --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-parse_kernel|radix_scatter_kernel|rowsort_kernel|fit_kernel}" -s ${NCU_S:-4} -c ${NCU_C:-4} -f -o gpurun_out/prof $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/ncu.log; ls -la gpurun_out/
