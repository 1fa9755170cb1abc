#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/host.txt
timeout 900 python bench.py --config C2_small --steps 2 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small exit $?"
tail -3 gpurun_out/bench_small.err
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 exit $?"
tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_small.json","gpurun_out/bench_c2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e lines/s  ms/step %.2f  e2e %s  cpu %s"%(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value")))
        print(" stage ms:", {k: round(v,3) for k,v in d["stage_ms_last_step"].items()})
        for r in d.get("kernels",[])[:14]: print("  %-24s n/step %5.1f  ms/step %8.3f  GB/s %s"%(r["kernel"], r["launches_per_step"], r["ms_per_step"], None if r["GBps"] is None else round(r["GBps"],1)))
        print(" roofline", d.get("roofline")); print(" path", d.get("path_roofline")); print(" clocks", d.get("clocks")); print(" counts", d.get("counts"))
    except Exception as e: print(f, "ERR", e)
PY
