#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -15 gpurun_out/s2_pytest.log
for shape in "56 64 64 3 1 1" "56 64 128 3 1 2" "28 128 128 3 1 1" "28 128 256 3 1 2" "14 256 256 3 1 1" "14 256 512 3 1 2" "7 512 256 3 1 1" "7 256 128 3 1 1" "14 256 128 14 0 1"; do
  python tools/conv_layer_run.py $shape 69824 5 >> gpurun_out/s2_layers.txt 2>&1
done
cat gpurun_out/s2_layers.txt
python tools/conv_bias_probe.py > gpurun_out/s2_bias.txt 2>&1; cat gpurun_out/s2_bias.txt
python tests/precision_report.py 4 2 450 fp32,fp32@64,fp32-tf32,fp32-simt > gpurun_out/s2_precision.txt 2>&1; cat gpurun_out/s2_precision.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/s2_bench.json"))
print("value %.0f e2e %.0f"%(d["value"],d["e2e"]["value"]), d["stage_ms_per_step"], d["parity_spot"]["per_tensor"], d["clocks"])
PY
python tools/conv_layer_run.py 14 256 512 3 1 2 69824 2 > gpurun_out/s2_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f -o gpurun_out/s2_conv256 python tools/conv_layer_run.py 14 256 512 3 1 2 69824 2 > gpurun_out/s2_ncu.log 2>&1
python tools/conv_layer_run.py 56 64 64 3 1 1 69824 2 > gpurun_out/s2_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f -o gpurun_out/s2_conv64 python tools/conv_layer_run.py 56 64 64 3 1 1 69824 2 > gpurun_out/s2_ncu.log 2>&1

timeout 300 python bench.py --no-cpu-baseline --sequences 1 --steps 50 > gpurun_out/s2_bench_b1.json 2> gpurun_out/s2_bench_b1.err; echo "b1 rc=$?"; tail -3 gpurun_out/s2_bench_b1.err
timeout 300 python bench.py --no-cpu-baseline --sequences 8 --steps 50 > gpurun_out/s2_bench_b8.json 2> gpurun_out/s2_bench_b8.err; echo "b8 rc=$?"; tail -3 gpurun_out/s2_bench_b8.err
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s2_bench_train.json 2> gpurun_out/s2_bench_train.err; echo "train rc=$?"; tail -3 gpurun_out/s2_bench_train.err
cat gpurun_out/s2_bench_b1.json gpurun_out/s2_bench_b8.json gpurun_out/s2_bench_train.json | cut -c1-1500
