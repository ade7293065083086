#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --scaling strong --no-cpu-baseline --no-parity-spot > gpurun_out/m${N}_strong.json 2> gpurun_out/m${N}_strong.err; echo "strong rc=$?"; tail -2 gpurun_out/m${N}_strong.err
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --scaling strong --graph --no-cpu-baseline --no-parity-spot > gpurun_out/m${N}_strong_graph.json 2> gpurun_out/m${N}_strong_graph.err; echo "strong graph rc=$?"; tail -2 gpurun_out/m${N}_strong_graph.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload train > gpurun_out/m${N}_train.json 2> gpurun_out/m${N}_train.err; echo "train rc=$?"; tail -2 gpurun_out/m${N}_train.err
python - <<PY
import json
for f in ("strong","strong_graph","train"):
    try:
        d=json.load(open("gpurun_out/m${N}_%s.json"%f))
        print(f, "value %.1f e2e %.1f ms/step %.2f n_gpus %d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["n_gpus"]), d.get("scaling"), d.get("clocks"))
    except Exception as e: print(f,"ERR",e)
PY
