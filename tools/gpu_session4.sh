#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
tail -8 gpurun_out/s4_pytest.log; grep -n "engine parity at the bench shape" gpurun_out/s4_pytest.log
POF_PROBE_TF32=1 python tools/conv_bias_probe.py > gpurun_out/s4_bias_tf32.txt 2>&1; cat gpurun_out/s4_bias_tf32.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s4_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-parity-spot --seq-chunk 128 > gpurun_out/s4_bench_c128.json 2> gpurun_out/s4_bench_c128.err; echo "bench rc=$?"; tail -3 gpurun_out/s4_bench_c128.err
python - <<'PY'
import json
for f in ("s4_bench","s4_bench_c128"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f,"value %.0f e2e %.0f"%(d["value"],d["e2e"]["value"]), d["stage_ms_per_step"], (d.get("parity_spot") or {}).get("per_tensor"), d["clocks"])
PY
