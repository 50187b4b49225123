timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "large_grids or integrate_parity" 2>&1 | tail -3
bash tools/_run2.sh
