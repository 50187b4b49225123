import json, os, subprocess, sys
for dbg in (0, 2, 2+4, 2+8, 2+4+8, 2+4+8+16, 1+2+4+8+16):
    env = dict(os.environ, SOMAX_B200_DEBUG=str(dbg))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    ks = {k["kernel"]: round(k["total_ms"] / k["launches"], 3) for k in d["roofline"]["kernels"] if k["kernel"].startswith("thomas")}
    print("dbg", dbg, ks, flush=True)
