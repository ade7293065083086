#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 4 11; do
  POF_TRAIN_SCANS_PER_CALL=$g timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s10_train_g$g.json 2> gpurun_out/s10_train_g$g.err; echo "g=$g rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/s10_train_g$g.json")); print("scans per call $g: ms/step %.2f value %.1f e2e %.1f loss %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["last_loss"]))
PY
done
timeout 900 python -m pytest tests/test_gpu_entrypoints.py tests/test_gpu_parity.py -m gpu -q -k "entry or drow_format or train or bn_act" > gpurun_out/s10_pytest.log 2>&1; tail -5 gpurun_out/s10_pytest.log
