"""Time YAML -> engine-ready arrays: the reference's loader + factor-graph build + per-assignment
tabulation against pydcop_b200.ingest + layout packing, on the same generated files.

Needs /root/reference (build container only).  Writes profiles/r01_ingest_timing.json.

    python tools/ingest_timing.py [n_vars ...]
"""
import itertools
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]


def write_yaml(path, n_vars, degree, d, seed, extensional):
    """Random binary soft-colouring problem in pyDcop's YAML form: intentional
    `c if a == b else 0`-style penalties, or dense extensional tables."""
    rng = np.random.default_rng(seed)
    n_con = n_vars * degree // 2
    with open(path, "w") as f:
        f.write(f"name: timing_{n_vars}\nobjective: min\n\ndomains:\n  colours:\n    values: "
                f"[{', '.join(str(i) for i in range(d))}]\n\nvariables:\n")
        for i in range(n_vars):
            f.write(f"  v{i}:\n    domain: colours\n")
        f.write("\nconstraints:\n")
        for j in range(n_con):
            a, b = rng.choice(n_vars, 2, replace=False)
            if extensional:
                t = rng.integers(0, 10, (d, d))
                by_cost = {}
                for x, y in itertools.product(range(d), range(d)):
                    by_cost.setdefault(int(t[x, y]), []).append(f"{x} {y}")
                f.write(f"  c{j}:\n    type: extensional\n    variables: [v{a}, v{b}]\n    values:\n")
                for cost, asg in by_cost.items():
                    f.write(f"      {cost}: {' | '.join(asg)}\n")
            else:
                w = int(rng.integers(1, 10))
                f.write(f"  c{j}:\n    type: intention\n    function: {w} if v{a} == v{b} else abs(v{a} - v{b}) * 0.1\n")
        f.write("\nagents: [a0]\n")
    return n_con


def time_reference(path):
    import ref_shim
    ref_shim.install()
    from pydcop.computations_graph import factor_graph
    from pydcop.dcop.yamldcop import load_dcop_from_file
    t0 = time.perf_counter()
    dcop = load_dcop_from_file([path])
    t1 = time.perf_counter()
    graph = factor_graph.build_computation_graph(dcop)
    t2 = time.perf_counter()
    n = 0
    for c in dcop.constraints.values():  # what a factor computation evaluates before cycle 1
        doms = [list(v.domain) for v in c.dimensions]
        for combo in itertools.product(*doms):
            c(**{v.name: x for v, x in zip(c.dimensions, combo)})
            n += 1
    t3 = time.perf_counter()
    return {"load_dcop_s": t1 - t0, "build_computation_graph_s": t2 - t1, "tabulate_s": t3 - t2,
            "total_s": t3 - t0, "table_entries": n, "nodes": len(graph.nodes)}


def time_ingest(path):
    from pydcop_b200 import ingest
    from pydcop_b200.layout import build_layout
    t0 = time.perf_counter()
    d = ingest.load_yaml(path)
    t1 = time.perf_counter()
    build_layout(**d.instance())
    t2 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "i.fgb")
        ingest.save_instance(p, d, names=False)
        t3 = time.perf_counter()
        e = ingest.load_instance(p)
        build_layout(**e.instance())
        t4 = time.perf_counter()
        size = os.path.getsize(p)
    return {"load_yaml_s": t1 - t0, "build_layout_s": t2 - t1, "total_s": t2 - t0,
            "tabulation": d.meta["tabulation"], "binary_load_and_layout_s": t4 - t3,
            "binary_bytes": size, "yaml_bytes": os.path.getsize(path)}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [200, 1000, 4000]
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        for n in sizes:
            for ext in (False, True):
                path = os.path.join(tmp, f"t{n}_{int(ext)}.yaml")
                n_con = write_yaml(path, n, 4, 10, seed=n, extensional=ext)
                mine = time_ingest(path)
                ref = time_reference(path)
                rows.append({"n_vars": n, "n_constraints": n_con, "domain": 10,
                             "form": "extensional" if ext else "intention",
                             "reference": ref, "pydcop_b200": mine,
                             "speedup": ref["total_s"] / mine["total_s"]})
                print(json.dumps(rows[-1]))
    out = os.path.join(ROOT, "profiles", "r01_ingest_timing.json")
    with open(out, "w") as f:
        json.dump({"host": "build container CPU, 1 thread", "rows": rows}, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
