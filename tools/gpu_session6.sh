#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
tail -8 gpurun_out/s6_pytest.log; grep -n "^E  " gpurun_out/s6_pytest.log | head -20
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s6_train.json 2> gpurun_out/s6_train.err; echo "train rc=$?"; tail -3 gpurun_out/s6_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s6_train.json")); print("train value %.1f ms/step %.2f e2e %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["roofline"]["frac"], d["last_loss"])
PY
python tools/tune_cutout.py > gpurun_out/s6_cutout_sweep.txt 2>&1; grep "S=11\|B= 256" gpurun_out/s6_cutout_sweep.txt
# ncu captures of the two HBM-bound kernels at the bench's launch shapes
python tools/gate_run.py 64 3 1 > gpurun_out/s6_gate_plain.log 2>&1 && cat gpurun_out/s6_gate_plain.log && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gate_stream_kernel -s 1 -c 1 -f -o gpurun_out/s6_gate python tools/gate_run.py 64 3 1 > gpurun_out/s6_gate_ncu.log 2>&1
python tools/gate_run.py 64 3 0 >> gpurun_out/s6_gate_plain.log 2>&1; tail -1 gpurun_out/s6_gate_plain.log
TUNE_ONLY=1 python tools/tune_cutout.py > gpurun_out/s6_cut_plain.log 2>&1 && \
TUNE_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cutout_scan_kernel -s 3 -c 1 -f -o gpurun_out/s6_cutout_scan python tools/tune_cutout.py > gpurun_out/s6_cut_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
