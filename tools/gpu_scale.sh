#!/bin/bash
# multi-GPU bench: N ranks on one box (gpurun --gpus N)
N=${N:-2}
mkdir -p gpurun_out
timeout ${TMO:-1500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-3} --warmup ${WARMUP:-3} ${BENCH_ARGS} > gpurun_out/bench_n$N${TAG}.json 2> gpurun_out/bench_n$N${TAG}.err
echo "n$N exit $?"; grep -v "^\[W\|Warning\|warn" gpurun_out/bench_n$N${TAG}.err | tail -${ERRTAIL:-12}
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N${TAG}.json").read().strip().splitlines()[-1])
    print("N=%d value %.3e lines/s  ms/step %.2f  non-kernel %.2f ms  host enqueue %.2f ms"%(d["n_gpus"], d["value"], d["ms_per_step"], d.get("non_kernel_ms_per_step",-1), d.get("host_enqueue_ms_per_step",-1)))
    print(" workload:", d["config"]["workload"])
    print(" stage ms:", {k: round(v,3) for k,v in d["stage_ms_last_step"].items()})
    for r in d.get("kernels",[])[:22]: print("  %-24s n/step %5.1f  ms/step %8.3f"%(r["kernel"], r["launches_per_step"], r["ms_per_step"]))
    print(" path", d.get("path_roofline")); print(" parity", d.get("parity")); print(" e2e", d.get("e2e")); print(" clocks", d.get("clocks"))
except Exception as e: print("ERR", e)
PY
