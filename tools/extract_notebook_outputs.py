"""Extract the numbers the REFERENCE ITSELF printed when its tutorials were executed.

The reference ships executed notebooks next to their jupytext sources
(/root/reference/content/tutorials/step09_laplace_2d.ipynb, step10_poisson_2d.ipynb,
step11_helmholtz_2d.ipynb, step17_shallow_water_2d.ipynb).  Their stored cell outputs are the only
reference-produced values for this path that exist anywhere (the reference cannot be run in this
image: no jax).  This script copies the stdout of the cells that exercise the hot path's
operators into tests/golden/reference_notebook_outputs.json, together with the parsed numbers.

    python tools/extract_notebook_outputs.py [/root/reference]

Only printed OUTPUTS are stored (a few hundred bytes), no reference source.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CELLS = {
    "step09_laplace_2d.ipynb": [7, 11, 13],
    "step10_poisson_2d.ipynb": [4, 12],
    "step11_helmholtz_2d.ipynb": [6, 12],
    "step17_shallow_water_2d.ipynb": [7, 9, 11, 18],
}
NUM = r"[-+]?\d+\.\d*(?:[eE][-+]?\d+)?"


def cell_text(cell):
    out = []
    for o in cell.get("outputs", []):
        if o.get("output_type") == "stream":
            out.append("".join(o["text"]))
        elif "text/plain" in o.get("data", {}) and "image/png" not in o.get("data", {}):
            out.append("".join(o["data"]["text/plain"]))
    return "".join(out)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    raw = {}
    for nb, cells in CELLS.items():
        d = json.load(open(os.path.join(ref, "content", "tutorials", nb)))
        for c in cells:
            raw[f"{nb}#cell{c}"] = cell_text(d["cells"][c])
    f = lambda pat, key: float(re.search(pat, raw[key]).group(1))  # noqa: E731
    s9, s11, s17 = "step09_laplace_2d.ipynb", "step11_helmholtz_2d.ipynb", "step17_shallow_water_2d.ipynb"
    conv = {int(n): float(e) for n, e in re.findall(r"n=\s*(\d+)\s+L2 error = (" + NUM + ")", raw[f"{s9}#cell13"])}
    helm = {float(l): float(a) for l, a in re.findall(r"lambda =\s*(" + NUM + r")\s+max\|phi\| = (" + NUM + ")", raw[f"{s11}#cell6"])}
    tail = lambda var: [float(x) for x in re.search(  # noqa: E731
        r"^\s+" + var + r"\s+\(time, y, x\).*?\.\.\.\s+(.*)$", raw[f"{s17}#cell11"], re.M).group(1).split()]
    parsed = {
        # PoissonSolver2D(bc="dirichlet"), rhs = -2 pi^2 sin(pi x) sin(pi y) on x = arange(Nx) * dx
        "poisson_dirichlet_l2_by_n": conv,
        "poisson_dirichlet_l2_n64": f(r"L2\s+error \(interior\): (" + NUM + ")", f"{s9}#cell7"),
        "poisson_dirichlet_linf_n64": f(r"Linf error \(interior\): (" + NUM + ")", f"{s9}#cell7"),
        "poisson_dirichlet_mode23_l2_n64": f(r"Higher-mode L2 error: (" + NUM + ")", f"{s9}#cell11"),
        # HelmholtzSolver2D(lambda_), rhs = sin(pi x) sin(pi y), n = 64
        "helmholtz_maxabs_by_lambda": helm,
        # NonlinearShallowWater2D, 32^2, wall, wind spin-up to t = 5e6 s (fp32)
        "swm17_energy_printed": f(r"Final KE: (" + NUM + ")", f"{s17}#cell9"),
        "swm17_max_abs_u": f(r"Max \|u\|: (" + NUM + ")", f"{s17}#cell9"),
        "swm17_max_abs_v": f(r"Max \|v\|: (" + NUM + ")", f"{s17}#cell9"),
        "swm17_tail_u": tail("u"), "swm17_tail_v": tail("v"), "swm17_tail_eta": tail("eta"),
        # eqx.filter_grad of sum(u^2) after t1 = 1e5 s
        "swm17_dloss_dviscosity": f(r"d\(loss\)/d\(viscosity\)\s+= (" + NUM + ")", f"{s17}#cell18"),
        "swm17_dloss_dwind": f(r"d\(loss\)/d\(wind_amplitude\) = (" + NUM + ")", f"{s17}#cell18"),
    }
    out = {"source": "stored cell outputs of /root/reference/content/tutorials/*.ipynb (somax v0.0.6 tree)",
           "made_by": "tools/extract_notebook_outputs.py", "raw_stdout": raw, "parsed": parsed}
    path = os.path.join(ROOT, "tests", "golden", "reference_notebook_outputs.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(parsed, indent=1))


if __name__ == "__main__":
    main()
