mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
timeout 600 python tools/tune_cutout.py > gpurun_out/cutout_sweep.txt 2>&1; grep "4096\|512" gpurun_out/cutout_sweep.txt | grep jrdb
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline_cutout']['frac'], d['roofline_cutout']['exact_arithmetic'], d['parity_spot'], d['gpu_launches'], d['stage_ms_per_step'])
PY
TUNE_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cutout_scan_exact -c 1 -f -o gpurun_out/exact_scan3 python tools/tune_cutout.py > gpurun_out/ncu_exact.log 2>&1; tail -2 gpurun_out/ncu_exact.log
