#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload train > gpurun_out/mt${N}_train.json 2> gpurun_out/mt${N}_train.err; echo "train rc=$?"; tail -2 gpurun_out/mt${N}_train.err
python - <<PY
import json
d=json.load(open("gpurun_out/mt${N}_train.json")); print("train N=$N value %.1f e2e %.1f ms/step %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]), d.get("clocks"))
PY
