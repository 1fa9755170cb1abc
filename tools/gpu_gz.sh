#!/bin/bash
# inflate tests + the gz e2e leg
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "inflate or compressed or gzip or dendro_working" 2>&1 | tail -6
timeout 500 python bench.py --steps 3 --warmup 1 --no-cpu --no-parity --e2e-input gz > gpurun_out/gz2.json 2> gpurun_out/gz2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/gz2.json").read().strip().splitlines()[-1]); e=d.get("e2e"); print("e2e gz ms", e.get("ms_per_step"), "text ms", e.get("plain_text_input",{}).get("ms_per_step"), e.get("inflate"))
except Exception as ex: print("ERR", ex); print(open("gpurun_out/gz2.err").read()[-1500:])
PY
