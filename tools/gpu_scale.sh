#!/bin/bash
# weak-scaling bench at N GPUs of one box: bash tools/gpu_scale.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "N=$N exit $?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.3e ms/step %.2f e2e %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("e2e")))
print(d["config"]["workload"]); print(d["stage_ms_last_step"]); print(d["clocks"])
PY
