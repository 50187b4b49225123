import json, os, subprocess, sys
dbgs = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 1 + 2]
for dbg in dbgs:
    env = dict(os.environ, SOMAX_B200_DEBUG=str(dbg))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    ks = {k["kernel"]: round(k["total_ms"] / k["launches"], 3) for k in d["roofline"]["kernels"] if k["kernel"].startswith(("thomas", "border"))}
    print("dbg", dbg, d["ms_per_step"], dict(sorted(ks.items())), flush=True)
