#!/bin/bash
# what the driver runs at round end, plus the evidence files: GPU tests, smoke, the default bench line, the reference arm, then the ncu launch
# list and the full capture of the top kernels (tools/gpu_final.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.4e ms/step %.3f e2e %s parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("parity",{}).get("ok")))
print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
r=json.loads(open("gpurun_out/bench_reference.json").read().strip().splitlines()[-1]); print("reference arm %.4e %s" % (r["value"], r["config"]["workload"]))
PY
bash tools/gpu_final.sh
# MCL on the device: whole process on a device-resident graph
timeout 300 python tools/mcl_timing.py 1.2 C2_small > gpurun_out/mcl_timing.jsonl 2> gpurun_out/mcl_timing.err; echo "mcl timing exit $?"; cat gpurun_out/mcl_timing.jsonl
