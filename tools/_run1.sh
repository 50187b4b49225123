timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "full_size" 2>&1 | tail -15
timeout 600 python tools/thomas_probe.py 0
