#!/bin/bash
# round-2 GPU session 1: tests at HEAD, bench, chunk sweep, narrow-layer ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench rc=$?"
for c in 8 16 32; do
  timeout 300 python bench.py --no-cpu-baseline --no-parity-spot --seq-chunk $c --steps 4 > gpurun_out/s1_bench_chunk$c.json 2> gpurun_out/s1_bench_chunk$c.err
done
for shape in "56 64 64 3 1 1" "56 64 128 3 1 2" "28 128 128 3 1 1" "28 128 256 3 1 2" "14 256 256 3 1 1" "14 256 512 3 1 2" "7 512 256 3 1 1" "7 256 128 3 1 1" "14 256 128 14 0 1"; do
  python tools/conv_layer_run.py $shape 69824 5 >> gpurun_out/s1_layers.txt 2>&1
done
cat gpurun_out/s1_layers.txt
n=0
for shape in "56 64 64 3 1 1" "28 128 128 3 1 1" "28 128 256 3 1 2" "14 256 512 3 1 2"; do
  n=$((n+1))
  python tools/conv_layer_run.py $shape 69824 2 > gpurun_out/s1_plain$n.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f -o gpurun_out/s1_conv$n python tools/conv_layer_run.py $shape 69824 2 > gpurun_out/s1_ncu$n.log 2>&1
done
ls -la gpurun_out | tail -20
