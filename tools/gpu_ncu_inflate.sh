#!/bin/bash
# ncu capture of inflate_kernel on the C2 workload fed as .gz (bench.py --e2e-input gz)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-input gz ${BENCH_ARGS}"
$CMD > gpurun_out/plain_gz.json 2> gpurun_out/plain_gz.err || { echo "plain run failed"; tail -5 gpurun_out/plain_gz.err; exit 1; }
python - <<PY
import json
d=json.loads(open("gpurun_out/plain_gz.json").read().strip().splitlines()[-1]); print("e2e", d.get("e2e"))
PY
ncu --set full --clock-control none --import-source on -k regex:"^inflate_kernel" -s 4 -c 1 -f -o gpurun_out/src_inflate $CMD > gpurun_out/ncu_inflate.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_inflate.log
ls -la gpurun_out/*.ncu-rep
