#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -12 gpurun_out/s3_pytest.log
python tools/conv_bias_probe.py > gpurun_out/s3_bias.txt 2>&1; cat gpurun_out/s3_bias.txt
python tests/precision_report.py 4 2 450 fp32,fp32@64 > gpurun_out/s3_precision.txt 2>&1; cat gpurun_out/s3_precision.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; tail -2 gpurun_out/s3_smoke.log
timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s3_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/s3_bench_ref.json 2> gpurun_out/s3_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/s3_bench_ref.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/s3_bench.json"))
print("value %.0f e2e %.0f"%(d["value"],d["e2e"]["value"]), d["stage_ms_per_step"], d["parity_spot"], d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["torch_gpu_baseline"]["value"])
PY
