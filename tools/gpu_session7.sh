#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -5 gpurun_out/s7_pytest.log; grep -n "^E  " gpurun_out/s7_pytest.log | head -10
timeout 300 python bench.py --workload train --steps 4 > gpurun_out/s7_train_plain.json 2> gpurun_out/s7_train_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4000 --csv --log-file gpurun_out/s7_train_launches.csv python bench.py --workload train --steps 4 > gpurun_out/s7_train_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s7_train_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4000: python bench.py --workload train --steps 4   [fused BN + LeakyReLU + pool operator]" 30 > gpurun_out/s7_train_launch_summary.txt; head -34 gpurun_out/s7_train_launch_summary.txt | cut -c1-170
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-spot > gpurun_out/s7_plain.json 2> gpurun_out/s7_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s7_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-spot > gpurun_out/s7_bench_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s7_bench_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 700: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-spot  (first 700 launches: the warm-up step and the two timed steps of the device-timed run)" 30 > gpurun_out/s7_bench_launch_summary.txt; head -24 gpurun_out/s7_bench_launch_summary.txt | cut -c1-170
