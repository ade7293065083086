set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cutout" > gpurun_out/exact_tests.log 2>&1; echo "rc=$?" >> gpurun_out/exact_tests.log
tail -15 gpurun_out/exact_tests.log
TUNE_ONLY=1 timeout 300 python tools/tune_cutout.py > gpurun_out/exact_sweep1.txt 2>&1; cat gpurun_out/exact_sweep1.txt
TUNE_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cutout_scan_exact -c 1 -o gpurun_out/exact_scan python tools/tune_cutout.py > gpurun_out/ncu_exact.log 2>&1; tail -3 gpurun_out/ncu_exact.log
