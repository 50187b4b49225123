timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -6
bash tools/_run2.sh
