#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
tail -5 gpurun_out/s9_pytest.log; grep -n "^E  " gpurun_out/s9_pytest.log | head -10
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s9_train.json 2> gpurun_out/s9_train.err; echo "train rc=$?"; tail -3 gpurun_out/s9_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s9_train.json")); print("train value %.1f ms/step %.2f e2e %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["roofline"]["frac"], d["last_loss"])
PY
timeout 300 python bench.py --workload train --steps 4 > gpurun_out/s9_train_plain.json 2> gpurun_out/s9_train_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file gpurun_out/s9_train_launches.csv python bench.py --workload train --steps 4 > gpurun_out/s9_train_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s9_train_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500: python bench.py --workload train --steps 4   [all scans through conv blocks 1-2 in one pass, per-scan batch statistics]" 24 > gpurun_out/s9_train_launch_summary.txt; head -28 gpurun_out/s9_train_launch_summary.txt | cut -c1-150
