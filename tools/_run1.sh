timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "large_grids" 2>&1 | tail -4
bash tools/_run2.sh
