python - <<'PY'
import json,subprocess,sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
d=json.loads(out)
print(d["ms_per_step"])
for k in d["roofline"]["kernels"][:4]: print(k)
PY
