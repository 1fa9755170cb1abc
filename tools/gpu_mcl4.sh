#!/bin/bash
# final check of the round: smoke (incl. MCL), a bench line with the opt-in MCL block on a small config, one ncu capture of the
# row-indexed expansion kernel
mkdir -p gpurun_out
timeout 70 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 80 python bench.py --config C2_small --steps 3 --warmup 3 --mcl 1.2 > gpurun_out/bench_C2_small_mcl.json 2> gpurun_out/bench_C2_small_mcl.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_C2_small_mcl.json").read().strip().splitlines()[-1])
    print("value %.3e ms/step %.3f parity %s e2e %s" % (d["value"], d["ms_per_step"], d.get("parity",{}).get("ok"), d.get("e2e",{}).get("value")))
    print("mcl", d.get("mcl"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_C2_small_mcl.err").read()[-600:])
PY
timeout 45 ncu --set full --clock-control none --import-source on -k regex:"mcl_expand_direct_kernel" -s 5 -c 2 -f -o gpurun_out/mcl_direct_prof python tools/mcl_timing.py 1.2 C2_tiny > gpurun_out/ncu_mcl_direct.log 2>&1; echo "ncu exit $?"
