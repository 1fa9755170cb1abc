#!/bin/bash
# parser tests + C2 step (twice) + a 128-species job with 9-character sequence ids
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "long_sequence or fuzz or boundaries or raw_matrices" 2>&1 | tail -2
for i in 1 2; do
  timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/p1.json 2> gpurun_out/p1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/p1.json").read().strip().splitlines()[-1]); print("C2 ms/step", d["ms_per_step"], [ (k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:2]])
PY
done
timeout 400 python bench.py --config C4 --species 128 --genes 5000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c4l.json 2> gpurun_out/c4l.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c4l.json").read().strip().splitlines()[-1]); print("128 species ms/step", d["ms_per_step"], [ (k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:2]])
PY
