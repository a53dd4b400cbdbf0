#!/usr/bin/env python
"""Random pyDcop YAML problems with random intentional constraints (arithmetic, abs, round, max, //, %,
**, comparisons, conditional expressions over int / float / range domains): the tables pydcop_b200.ingest
produces against the UNMODIFIED reference loader evaluated assignment by assignment.

TEST INFRASTRUCTURE (build container only).   python oracle/fuzz_ingest_vs_reference.py [n_files]
Last line: "constraints compared N bad B {tabulation statistics}".
"""
import itertools
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]
import ref_shim; ref_shim.install()
import logging; logging.disable(logging.CRITICAL)
import numpy as np
from pydcop.dcop.yamldcop import load_dcop
from pydcop_b200 import ingest

def rand_expr(rnd, names, depth=0):
    if depth > 2 or rnd.random() < 0.3:
        return rnd.choice(names + [str(rnd.choice([0, 1, 2, 3, 0.5, -1.5, 7]))])
    k = rnd.random()
    a, b = rand_expr(rnd, names, depth+1), rand_expr(rnd, names, depth+1)
    if k < 0.45: return f"({a} {rnd.choice(['+','-','*'])} {b})"
    if k < 0.55: return f"abs({a})"
    if k < 0.65: return f"({a} if {b} {rnd.choice(['==','<','>=','!='])} {rand_expr(rnd,names,depth+1)} else {rand_expr(rnd,names,depth+1)})"
    if k < 0.72: return f"round({a})"
    if k < 0.79: return f"max({a}, {b})"
    if k < 0.86: return f"({a} // 2)"
    if k < 0.93: return f"({a} % 3)"
    return f"({a} ** 2)"

bad = 0; total = 0; stats = {"vectorised":0,"scalar":0,"shared":0}
for seed in range(int(sys.argv[1]) if len(sys.argv)>1 else 40):
    rnd = random.Random(seed)
    doms = {"di": [-2, -1, 0, 1, 2, 3], "df": [0.5, 1.25, -2.0, 3.0], "dr": ["0..4"]}
    lines = ["name: fuzz", "objective: min", "domains:"]
    for d, v in doms.items():
        lines.append(f"  {d}: {{values: {v}}}")
    lines.append("variables:")
    names = [f"v{i}" for i in range(5)]
    for n in names:
        lines.append(f"  {n}: {{domain: {rnd.choice(list(doms))}}}")
    lines.append("constraints:")
    for c in range(8):
        sub = rnd.sample(names, rnd.randint(1, 3))
        expr = rand_expr(rnd, sub)
        lines.append(f"  c{c}: {{type: intention, function: \"{expr}\"}}")
    var_dom = {n: None for n in names}
    for ln in lines:
        for n in names:
            if ln.startswith(f"  {n}: {{domain: "):
                var_dom[n] = ln.split("domain: ")[1].rstrip("}")
    real = {"di": doms["di"], "df": doms["df"], "dr": [0, 1, 2, 3, 4]}
    for c in range(3):      # extensional constraints: cost -> "a b | c d", some cells left to `default`
        sub = rnd.sample(names, rnd.randint(1, 2))
        cells = list(itertools.product(*[real[var_dom[n]] for n in sub]))
        by_cost = {}
        for cell in cells:
            if rnd.random() < 0.8:
                by_cost.setdefault(rnd.choice([0, 1, 2.5, -3, 10]), []).append(" ".join(str(x) for x in cell))
        lines.append(f"  e{c}:")
        lines.append("    type: extensional")
        lines.append(f"    variables: {sub[0] if len(sub) == 1 and rnd.random() < 0.5 else '[' + ', '.join(sub) + ']'}")
        lines.append(f"    default: {rnd.choice([0, 7, -1.5])}")
        lines.append("    values:")
        for cost, asg in by_cost.items():
            lines.append(f"      {cost}: \"{' | '.join(asg)}\"")
        if not by_cost:
            lines[-1] = "    values: {}"
    text = "\n".join(lines) + "\n"
    try:
        ref = load_dcop(text)
    except Exception as e:
        continue
    try:
        mine = ingest.loads_yaml(text)
    except Exception as e:
        # acceptable only if the reference would also fail when evaluating
        try:
            for c in ref.constraints.values():
                for combo in itertools.product(*[list(v.domain) for v in c.dimensions]):
                    c(**{v.name: x for v, x in zip(c.dimensions, combo)})
            print("MINE FAILED but reference evaluates:", seed, repr(e)); bad += 1
        except Exception:
            pass
        continue
    for k in stats: stats[k] += mine.meta["tabulation"][k]
    a = mine.arrays
    for ci, cn in enumerate(mine.con_names):
        c = ref.constraints[cn]
        scope = [mine.var_names[i] for i in a["edge_var"][a["factor_ptr"][ci]:a["factor_ptr"][ci+1]]]
        if sorted(scope) != sorted(v.name for v in c.dimensions):
            print("SCOPE", seed, cn, scope, [v.name for v in c.dimensions]); bad += 1; continue
        doms_ = [mine.values_of(n) for n in scope]
        t = a["tables"][a["table_off"][ci]:a["table_off"][ci+1]]
        want = [float(c(**dict(zip(scope, combo)))) for combo in itertools.product(*doms_)]
        total += 1
        if not np.array_equal(t, np.array(want), equal_nan=True):
            print("TABLE", seed, cn, getattr(c, "expression", "extensional"), t[:5], want[:5]); bad += 1
print("constraints compared", total, "bad", bad, stats)
