timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/b_full.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/b_full.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["step"]["frac"], d["clocks"])
PY
