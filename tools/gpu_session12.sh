#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gate or attention or training or spatial" > gpurun_out/s12_pytest_gate.log 2>&1; tail -4 gpurun_out/s12_pytest_gate.log; grep -n "^E  " gpurun_out/s12_pytest_gate.log | head
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s12_train.json 2> gpurun_out/s12_train.err; echo "train rc=$?"; tail -3 gpurun_out/s12_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s12_train.json")); print("train ms/step %.2f value %.1f e2e %.1f"%(d["ms_per_step"],d["value"],d["e2e"]["value"]), "gate bwd roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["last_loss"])
PY
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_pytest.log; tail -4 gpurun_out/s12_pytest.log
