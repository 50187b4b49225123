"""Write configs/simulation/*.yaml with exactly the VALUES of the reference's RunSpec files
(/root/reference/configs/simulation/*.yaml), so that a run with a shipped config is comparable with
the reference's DVC runs.  Values only (yaml.safe_load -> yaml.safe_dump), plus a header.
Run: python tools/sync_configs.py [/root/reference]"""
import os
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = os.path.join(ref, "configs", "simulation")
dst = os.path.join(ROOT, "configs", "simulation")
os.makedirs(dst, exist_ok=True)
for name in sorted(os.listdir(src)):
    if not name.endswith(".yaml"):
        continue
    spec = yaml.safe_load(open(os.path.join(src, name)))
    with open(os.path.join(dst, name), "w") as f:
        f.write(f"# RunSpec of the reference run `{name}` (same values as configs/simulation/{name} of\n"
                "# jejjohnson/somax v0.0.6; written by tools/sync_configs.py).  Optional extra key read by\n"
                "# somax_b200 only: testcase.grid.dtype (float32 | float64, default float32).\n")
        yaml.safe_dump(spec, f, sort_keys=False)
    print("wrote", os.path.join("configs", "simulation", name))
