#!/bin/bash
mkdir -p gpurun_out
for g in 3 4 6; do
  POF_TRAIN_SCANS_PER_CALL=$g timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s11_train_g$g.json 2> gpurun_out/s11_train_g$g.err; echo "g=$g rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/s11_train_g$g.json")); print("scans per call $g: ms/step %.2f value %.1f e2e %.1f loss %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["last_loss"]))
PY
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s11_pytest.log; tail -5 gpurun_out/s11_pytest.log; grep -n "^E  " gpurun_out/s11_pytest.log | head
timeout 300 python bench.py --workload train --steps 4 > gpurun_out/s11_train_plain.json 2> gpurun_out/s11_train_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 2000 --csv --log-file gpurun_out/s11_train_launches.csv python bench.py --workload train --steps 4 > gpurun_out/s11_train_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s11_train_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 2000: python bench.py --workload train --steps 4   [4 scans per call through conv blocks 1-2, per-scan batch statistics]" 24 > gpurun_out/s11_train_launch_summary.txt; head -22 gpurun_out/s11_train_launch_summary.txt | cut -c1-150
