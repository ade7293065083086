#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
tail -5 gpurun_out/s8_pytest.log; grep -n "^E  " gpurun_out/s8_pytest.log | head -10
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s8_train.json 2> gpurun_out/s8_train.err; echo "train rc=$?"; tail -3 gpurun_out/s8_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s8_train.json")); print("train value %.1f ms/step %.2f e2e %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["roofline"]["frac"], d["last_loss"])
PY
timeout 300 python bench.py --workload train --steps 4 > gpurun_out/s8_train_plain.json 2> gpurun_out/s8_train_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file gpurun_out/s8_train_launches.csv python bench.py --workload train --steps 4 > gpurun_out/s8_train_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s8_train_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000: python bench.py --workload train --steps 4   [fused BN + LeakyReLU + pool operator, first-layer kernels]" 24 > gpurun_out/s8_train_launch_summary.txt; head -28 gpurun_out/s8_train_launch_summary.txt | cut -c1-150
timeout 600 python bench.py > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s8_bench.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/s8_bench_ref.json 2> gpurun_out/s8_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/s8_bench.json"))
print("value %.0f e2e %.0f"%(d["value"],d["e2e"]["value"]), d["stage_ms_per_step"], d["parity_spot"]["per_tensor"], d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["torch_gpu_baseline"]["value"])
print(d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline_gate"]["frac"], d["roofline_gate"]["traffic"], d["roofline_gate"]["with_operand_split"]["frac"], d["roofline_cutout"]["frac"], d["roofline_cutout"]["traffic"])
r=json.load(open("gpurun_out/s8_bench_ref.json")); print("ref", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
PY
