mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cutout" 2>&1 | tail -3
timeout 600 python tools/tune_cutout.py 2>&1 | grep "EXACT " | grep "4096\|512"
