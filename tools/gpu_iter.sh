#!/bin/bash
# test + bench iteration
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -${PYTEST_TAIL:-15} gpurun_out/pytest_gpu.log
if [ -n "${SANITIZE}" ]; then
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${SANITIZE}" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitizer.log | tail -8
fi
if [ -z "${NO_BENCH}" ]; then
timeout 1500 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 exit $?"
tail -3 gpurun_out/bench_c2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e lines/s  ms/step %.2f  e2e %s  cpu %s"%(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value")))
        print(" stage ms:", {k: round(v,3) for k,v in d["stage_ms_last_step"].items()})
        for r in d.get("kernels",[])[:18]: print("  %-24s n/step %5.1f  ms/step %8.3f  GB/s %s"%(r["kernel"], r["launches_per_step"], r["ms_per_step"], None if r["GBps"] is None else round(r["GBps"],1)))
        print(" roofline", d.get("roofline")); print(" path", d.get("path_roofline")); print(" clocks", d.get("clocks")); print(" counts", d.get("counts")); print(" parity", d.get("parity")); print(" e2e", d.get("e2e"))
    except Exception as e: print(f, "ERR", e)
PY
fi
