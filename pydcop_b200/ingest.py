"""Fast ingestion: pyDcop YAML / pyDcop objects / a binary container -> the array front door.

SURVEY.md §8(f).2.  The reference turns a YAML file into Python objects (pydcop/dcop/yamldcop.py),
builds the factor graph with an O(|V|*|C|) scan (find_dependent_relations,
pydcop/dcop/relations.py:1245-1247 via computations_graph/factor_graph.py:277-280) and evaluates
intentional constraints one assignment at a time through ExpressionFunction
(pydcop/utils/expressionfunction.py:127-144).  None of that scales to the 10^5..10^6-variable
instances the engine is built for, so this module goes from the same YAML text straight to flat
arrays (`pydcop_b200.layout.build_layout` input):

  * `load_yaml` / `loads_yaml`   — the YAML grammar of yamldcop.py:99-122,140-152,167-197,215-280
    (domains incl. "a..b" ranges, variables with initial_value / cost_function / noise_level,
    intentional constraints incl. multi-line bodies and `source:` files, extensional constraints
    incl. `default` and "a b | c d" assignment lists);
  * `tabulate_expression`        — dense table of a Python expression, evaluated ONCE on broadcast
    numpy axes instead of once per assignment, checked against scalar evaluation on sampled
    entries and falling back to the scalar loop whenever numpy semantics could differ
    (`if`/`else`, `and`, `max(a, b)` … raise on arrays and take the fallback);
    identical expressions up to variable renaming over identical domains are tabulated once;
  * `from_dcop`                  — the same arrays from an already-built pyDcop DCOP object
    (duck-typed: no pydcop import);
  * `save_instance` / `load_instance` — binary container (header + 4 KiB-aligned raw arrays) that
    is memory-mapped on load: YAML's extensional form (yamldcop.py:240-277) is unusable at 10^6
    scale.

The graph arrays follow the reference's orders: variables and constraints in file order, scope
order = table axis order, every variable's incident edges in constraint order
(factor_graph.py:277-280).  For an intentional constraint the reference's scope order is the
iteration order of a Python set (expressionfunction.py:74,220 — it changes from one interpreter
run to the next); here it is the order of first appearance in the expression, which is one of
the orders the reference can produce.
"""
import ast
import builtins
import importlib.util
import itertools
import json
import os
import random
import re
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

try:  # the C loader is ~10x faster on large files; same grammar
    from yaml import CSafeLoader as _YamlLoader
except ImportError:  # pragma: no cover
    from yaml import SafeLoader as _YamlLoader
import yaml


class DcopFormatError(ValueError):
    """Malformed DCOP description (mirrors the ValueError / DcopInvalidFormatError the reference
    raises from yamldcop.py:102-105,172-176,221-226,279-283)."""


# --------------------------------------------------------------------------------------------
# instance container
# --------------------------------------------------------------------------------------------
ARRAY_KEYS = ("dom_size", "factor_ptr", "edge_var", "table_off", "tables", "unary", "var_ptr",
              "var_edge", "init_value")


class GeneratedNames(Sequence):
    """Names <prefix>0 .. <prefix>{n-1} without materialising n strings (array front door at 10^5..10^6 nodes)."""

    def __init__(self, prefix: str, n: int):
        self.prefix, self.n = prefix, int(n)

    def __len__(self):
        return self.n

    def __iter__(self):   # Sequence's default goes through __getitem__ (bounds / slice checks) once per item
        p = self.prefix
        return (f"{p}{k}" for k in range(self.n))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [f"{self.prefix}{k}" for k in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return f"{self.prefix}{i}"

    def index(self, name, *a):
        if isinstance(name, str) and name.startswith(self.prefix) and name[len(self.prefix):].isdigit():
            k = int(name[len(self.prefix):])
            if k < self.n and f"{self.prefix}{k}" == name:
                return k
        raise ValueError(name)

    def __eq__(self, other):
        return len(other) == self.n and all(a == b for a, b in zip(self, other))


@dataclass
class DcopArrays:
    """A DCOP as flat arrays plus the names needed to report a solution."""
    name: str
    objective: str                      # "min" | "max"
    arrays: Dict[str, np.ndarray]       # ARRAY_KEYS
    var_names: List[str]
    con_names: List[str]
    domain_values: Dict[str, list]      # domain name -> values
    var_domain: List[str]               # domain name of each variable
    meta: Dict[str, Any] = field(default_factory=dict)

    @property
    def n_vars(self):
        return len(self.var_names)

    @property
    def n_constraints(self):
        return len(self.con_names)

    def instance(self) -> Dict[str, np.ndarray]:
        """The dict `pydcop_b200.layout.layout_from_instance` takes."""
        return dict(self.arrays)

    def values_of(self, var: Union[int, str]) -> list:
        i = var if isinstance(var, int) else self.var_names.index(var)
        return self.domain_values[self.var_domain[i]]

    def assignment(self, value_index: Sequence[int]) -> Dict[str, Any]:
        """{variable name: domain value} from per-variable value indices (engine.values())."""
        idx = np.asarray(value_index).astype(np.int64).tolist()
        if len(self.domain_values) == 1:      # one shared domain: no per-variable lookup
            dom = next(iter(self.domain_values.values()))
            return {n: dom[k] for n, k in zip(self.var_names, idx)}
        doms = [self.domain_values[d] for d in self.var_domain]
        return {n: doms[i][k] for i, (n, k) in enumerate(zip(self.var_names, idx))}

    def cost(self, value_index: Sequence[int]) -> float:
        """Sum of constraint and variable costs of an assignment (DCOP.solution_cost,
        pydcop/dcop/dcop.py:319-367, without the infinity/violation split): host check for small
        cases; the engine's `solution_cost()` is the device version."""
        a = self.arrays
        idx = np.asarray(value_index, dtype=np.int64)
        fp, ev = a["factor_ptr"].astype(np.int64), a["edge_var"].astype(np.int64)
        dom = a["dom_size"].astype(np.int64)
        lin = np.zeros(len(fp) - 1, dtype=np.int64)
        arity = np.diff(fp)
        for j in range(int(arity.max(initial=0))):
            m = arity > j
            e = fp[:-1][m] + j
            lin[m] = lin[m] * dom[ev[e]] + idx[ev[e]]
        total = float(a["tables"][a["table_off"][:-1] + lin].astype(np.float64).sum())
        uoff = np.concatenate([[0], np.cumsum(dom)])[:-1]
        return total + float(a["unary"][uoff + idx].sum())


# --------------------------------------------------------------------------------------------
# expressions
# --------------------------------------------------------------------------------------------
_BUILTIN_NAMES = set(dir(builtins))


class _Names(ast.NodeVisitor):
    """Free names of an expression / function body: loaded, never stored, not imported, not a
    builtin, not `source*` (same rule as expressionfunction.py:178-205), in order of first
    appearance."""

    def __init__(self):
        self.loaded, self.stored, self.imported, self.has_return = {}, set(), set(), False

    def visit_Name(self, node):
        if isinstance(node.ctx, ast.Load):
            self.loaded.setdefault(node.id, (node.lineno, node.col_offset))
        elif isinstance(node.ctx, ast.Store):
            self.stored.add(node.id)

    def visit_Return(self, node):
        self.has_return = True
        self.generic_visit(node)

    def visit_Import(self, node):
        self.imported.update(n.name for n in node.names)

    visit_ImportFrom = visit_Import

    def free(self) -> List[str]:
        names = [n for n in self.loaded if n not in self.stored and n not in self.imported
                 and n not in _BUILTIN_NAMES and not n.startswith("source")]
        return sorted(names, key=lambda n: self.loaded[n])


class _Rename(ast.NodeTransformer):
    def __init__(self, mapping):
        self.mapping = mapping

    def visit_Name(self, node):
        if node.id in self.mapping:
            return ast.copy_location(ast.Name(id=self.mapping[node.id], ctx=node.ctx), node)
        return node


class Expression:
    """A Python expression (or function body with `return`) over named variables.

    Restates ExpressionFunction (expressionfunction.py:56-144): keyword-only call, free names are
    the arguments, a body containing `return` is wrapped as a function body, `source.<fn>` refers
    to the module loaded from `source_file`."""

    def __init__(self, expression: str, source_file: Optional[str] = None):
        self.text = expression.lstrip()
        self.source_file = str(source_file) if source_file is not None else None
        try:
            tree = ast.parse(self.text)
        except SyntaxError as e:
            raise SyntaxError(f"Syntax error in string expression: '{self.text}'") from e
        v = _Names()
        v.visit(tree)
        self.variable_names = v.free()
        self.has_return = v.has_return
        canon = _Rename({n: f"_a{i}" for i, n in enumerate(self.variable_names)}).visit(tree)
        #: identical for expressions that differ only by the names of their variables
        self.canonical_key = (ast.dump(canon), self.source_file)
        if self.has_return:
            body = self.text if self.text.startswith("\n") else "\n" + self.text
            src = f"def f({', '.join(self.variable_names)}):" + body.replace("\n", "\n    ")
        else:
            src = f"def f({', '.join(self.variable_names)}):\n    return {self.text}"
        g = {"__builtins__": builtins}
        if self.source_file is not None:
            spec = importlib.util.spec_from_file_location("source", self.source_file)
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            g["source"] = module
        local = {}
        try:
            exec(compile(src, "<dcop expression>", "exec"), g, local)
        except SyntaxError as e:
            raise SyntaxError(f"Syntax error in string expression: '{self.text}'") from e
        self.func = local["f"]

    def __call__(self, **kwargs):
        missing = set(self.variable_names) - set(kwargs)
        extra = set(kwargs) - set(self.variable_names)
        if missing:
            raise TypeError("Missing named argument(s) " + str(missing))
        if extra:
            raise TypeError("Unexpected argument(s) " + str(extra))
        return self.func(**kwargs)


def _axis_array(values: list):
    """Domain values as a numpy axis: int64 / float64 when all values are plain numbers (their
    arithmetic is then IEEE-identical to Python's for |ints| < 2^53), else an object array
    (element-wise Python semantics)."""
    if all(type(v) is int for v in values) and all(abs(v) < 2 ** 53 for v in values):
        return np.array(values, dtype=np.int64)
    if all(type(v) in (int, float) for v in values):
        return np.array(values, dtype=np.float64)
    a = np.empty(len(values), dtype=object)
    a[:] = values
    return a


def _tabulate_scalar(func, names, domains, shape):
    t = np.empty(shape, dtype=np.float64)
    flat = t.reshape(-1)
    for k, combo in enumerate(itertools.product(*domains)):
        flat[k] = func(**dict(zip(names, combo)))
    return t


def tabulate_expression(func, names: Sequence[str], domains: Sequence[list],
                        n_check: int = 24, stats: Optional[dict] = None) -> np.ndarray:
    """Dense float64 cost table of `func(**{name: value})`, axis i <-> names[i], row-major —
    the layout NAryMatrixRelation stores (relations.py:716-733).

    One vectorised evaluation on broadcast axes; accepted only if the result has a numeric dtype,
    broadcasts to the table shape and equals the scalar evaluation bit for bit — on EVERY entry of a
    table of up to 2048 entries, on `n_check` sampled entries plus the two corners of a larger one — and,
    when an axis holds integers, equals a second vectorised evaluation on float64 axes (numpy integers
    wrap silently on overflow where Python's grow: a wrapped `x ** k` cannot survive that comparison).
    Anything else — an exception, an object result, a mismatch — falls back to one scalar call per
    entry, i.e. what the reference does."""
    shape = tuple(len(d) for d in domains)
    size = int(np.prod(shape)) if shape else 1
    if not names:
        return np.full(shape, float(func()), dtype=np.float64)
    table = None
    if size > 4:
        axes = []
        for k, d in enumerate(domains):
            sh = [1] * len(shape)
            sh[k] = len(d)
            axes.append(_axis_array(list(d)).reshape(sh))
        try:
            with np.errstate(all="raise"):
                out = func(**dict(zip(names, axes)))
            out = np.asarray(out)
            if out.dtype.kind in "biuf":
                table = np.ascontiguousarray(np.broadcast_to(out, shape), dtype=np.float64)
                if any(a.dtype.kind in "iu" for a in axes):      # integer overflow guard
                    fout = np.asarray(func(**dict(zip(names, [a.astype(np.float64) if a.dtype.kind in "iu" else a
                                                               for a in axes]))))
                    ftab = np.broadcast_to(fout, shape).astype(np.float64)
                    if not np.array_equal(table, ftab, equal_nan=True):
                        table = None
        except Exception:  # noqa: BLE001 — any failure means "numpy semantics differ": fall back
            table = None
        if table is not None:
            rng = np.random.default_rng(size)
            picks = (range(size) if size <= 2048 else
                     {0, size - 1} | set(int(x) for x in rng.integers(0, size, min(n_check, size))))
            flat = table.reshape(-1)
            for p in picks:
                idx = np.unravel_index(p, shape)
                want = func(**{n: domains[k][i] for k, (n, i) in enumerate(zip(names, idx))})
                try:
                    want = float(want)
                except (TypeError, ValueError):
                    table = None
                    break
                if not (want == flat[p] or (want != want and flat[p] != flat[p])):
                    table = None
                    break
    if stats is not None:
        stats["vectorised" if table is not None else "scalar"] = \
            stats.get("vectorised" if table is not None else "scalar", 0) + 1
    if table is None:
        table = _tabulate_scalar(func, names, [list(d) for d in domains], shape)
    return table


# --------------------------------------------------------------------------------------------
# YAML
# --------------------------------------------------------------------------------------------
def _domain_values(spec) -> list:
    """yamldcop.py:140-152,479-501: a list of values, or one string "lo..hi" (inclusive ints)."""
    values = spec["values"]
    if len(values) == 1 and isinstance(values[0], str) and ".." in values[0]:
        s = values[0]
        try:
            i = s.index("..")
            return list(range(int(s[:i]), int(s[i + 2:]) + 1))
        except ValueError:
            vals = [v.strip() for v in s[1:].split(",")]
            try:
                return [int(v) for v in vals]
            except ValueError:
                return vals
    return list(values)


def _value_lookup(dom_vals: list) -> Dict[str, int]:
    """Domain.to_domain_value (objects.py:137-165) as a table: str(value) -> index of the FIRST
    value with that representation."""
    lut: Dict[str, int] = {}
    for i, v in enumerate(dom_vals):
        lut.setdefault(str(v), i)
    return lut


_IDENT = re.compile(r"(?<![\w.])[A-Za-z_]\w*")


class TableCache:
    """Tables of intentional constraints, shared between constraints that are the same function
    of same-domain variables up to renaming (graph colouring: |C| constraints, 1 table).

    Two levels.  `canonical()` renames the known variable names in the expression TEXT to
    positional placeholders with one regex pass, so that all constraints of one form share one
    parsed + compiled `Expression` (parsing and compiling dominate the load time otherwise);
    texts with string literals, where a textual rename could touch a literal, skip this level and
    are compared through their renamed AST instead (`Expression.canonical_key`)."""

    def __init__(self):
        self.tables: Dict[Any, np.ndarray] = {}
        self.expressions: Dict[Any, Expression] = {}
        self.stats = {"vectorised": 0, "scalar": 0, "shared": 0}

    def expression(self, text: str, source_file=None) -> Expression:
        key = (text, str(source_file) if source_file is not None else None)
        e = self.expressions.get(key)
        if e is None:
            e = self.expressions[key] = Expression(text, source_file)
        return e

    def canonical(self, text: str, known, source_file=None):
        """(shared Expression over placeholders _a0.., real variable names in placeholder order),
        or None when the text cannot be renamed safely."""
        if "'" in text or '"' in text:
            return None
        order: Dict[str, str] = {}
        clash = []

        def repl(m):
            n = m.group(0)
            if n.startswith("_a"):  # would be mistaken for a placeholder
                clash.append(n)
            if n not in known or n in _BUILTIN_NAMES or n.startswith("source"):
                return n
            ph = order.get(n)
            if ph is None:
                ph = order[n] = f"_a{len(order)}"
            return ph

        canon = _IDENT.sub(repl, text)
        if clash:
            return None
        return self.expression(canon, source_file), list(order)

    def get(self, expr: Expression, doms: List[Tuple[str, list]], names=None):
        key = (expr.canonical_key, tuple(d[0] for d in doms))
        t = self.tables.get(key)
        if t is None:
            t = tabulate_expression(expr.func, names or expr.variable_names, [d[1] for d in doms],
                                    stats=self.stats).reshape(-1)
            self.tables[key] = t
        else:
            self.stats["shared"] += 1
        return t


def loads_yaml(text: str, main_dir: Optional[str] = None, seed: Optional[int] = None) -> DcopArrays:
    """Parse a pyDcop YAML DCOP (yamldcop.py:99-122) into arrays.  Agents, routes, hosting costs
    and distribution hints are control-plane data the GPU path has no use for and are ignored.
    `seed` fixes the U(0, noise_level) draws of `noise_level` variables (objects.py:566-567;
    the reference draws from the global `random`)."""
    loaded = yaml.load(text, Loader=_YamlLoader)
    if not isinstance(loaded, dict) or "name" not in loaded:
        raise DcopFormatError("Missing name in dcop string")
    if loaded.get("objective") not in ("min", "max"):
        raise DcopFormatError("Objective is mandatory and must be min or max")

    domain_values = {n: _domain_values(d) for n, d in (loaded.get("domains") or {}).items()}
    if loaded.get("external_variables"):
        raise NotImplementedError("external_variables are not supported by the array ingestion "
                                  "(factor_graph.py:265 does not build computations for them either)")

    rnd = random.Random(seed) if seed is not None else random
    var_names, var_domain, init_value, unary = [], [], [], []
    for v_name, v in (loaded.get("variables") or {}).items():
        v = v or {}
        if v.get("domain") not in domain_values:
            raise DcopFormatError(f"unknown domain {v.get('domain')!r} for variable {v_name}")
        dvals = domain_values[v["domain"]]
        iv = v.get("initial_value")
        if iv and iv not in dvals:  # same truthiness test as yamldcop.py:172
            raise DcopFormatError(f"initial value {iv} is not in the domain {v['domain']} of the "
                                  f"variable {v_name}")
        var_names.append(str(v_name))
        var_domain.append(v["domain"])
        init_value.append(dvals.index(iv) if iv is not None and iv in dvals else -1)
        if "cost_function" in v:
            e = Expression(str(v["cost_function"]))
            if e.variable_names != [str(v_name)]:  # objects.py:485-494
                raise DcopFormatError(f"cost function of {v_name} must depend on {v_name} only, "
                                      f"not {e.variable_names}")
            costs = [float(e.func(**{str(v_name): x})) for x in dvals]
            if "noise_level" in v:
                costs = [c + rnd.uniform(0, v["noise_level"]) for c in costs]
            unary.extend(costs)
        else:
            unary.extend([0.0] * len(dvals))
    vidx = {n: i for i, n in enumerate(var_names)}
    dom_size = np.array([len(domain_values[d]) for d in var_domain], dtype=np.int32)

    con_names, factor_ptr, edge_var, tables = [], [0], [], []
    cache, luts = TableCache(), {}
    for c_name, c in (loaded.get("constraints") or {}).items():
        ctype = (c or {}).get("type")
        if ctype == "intention":
            src = c.get("source")
            if src is not None and not os.path.isabs(src):
                src = os.path.join(str(main_dir) if main_dir is not None else ".", src)
            text = str(c["function"])
            shared = cache.canonical(text, vidx, src)
            if shared is not None:
                e, real = shared
                # placeholders the expression really depends on, in order of first appearance
                try:
                    scope_names = [real[int(n[2:])] if n.startswith("_a") else None
                                   for n in e.variable_names]
                except (ValueError, IndexError):
                    scope_names = [None]
                missing = [n for n, r in zip(e.variable_names, scope_names) if r is None]
            else:
                e = cache.expression(text, src)
                scope_names = list(e.variable_names)
                missing = [n for n in scope_names if n not in vidx]
            if missing:  # relations.py:1301-1305
                raise DcopFormatError(f'Missing variable {missing[0]} for string-based function '
                                      f'"{text.lstrip()}"')
            scope = [vidx[n] for n in scope_names]
            doms = [(var_domain[i], domain_values[var_domain[i]]) for i in scope]
            table = cache.get(e, doms)
        elif ctype == "extensional":
            scope_names = c["variables"]
            if not isinstance(scope_names, list):  # single-variable form, yamldcop.py:230-243
                scope_names = [str(scope_names).strip()]
            for n in scope_names:
                if n not in vidx:
                    raise DcopFormatError(f"unknown variable {n} in constraint {c_name}")
            scope = [vidx[n] for n in scope_names]
            table = _extensional_table(c_name, c, scope, var_domain, domain_values, luts)
        else:
            raise DcopFormatError(f"Error in constraint {c_name} definition: type is mandatory "
                                  'and must be "intention" or "extensional"')
        con_names.append(str(c_name))
        edge_var.extend(scope)
        factor_ptr.append(len(edge_var))
        tables.append(table)

    arrays = _assemble(dom_size, factor_ptr, edge_var, tables, unary, init_value)
    return DcopArrays(name=str(loaded["name"]), objective=loaded["objective"], arrays=arrays,
                      var_names=var_names, con_names=con_names, domain_values=domain_values,
                      var_domain=var_domain,
                      meta={"description": loaded.get("description", ""),
                            "tabulation": dict(cache.stats)})


def _extensional_table(c_name, c, scope, var_domain, domain_values, luts):
    """yamldcop.py:227-277: {cost: "v1 v2 | v1 v2 …"} (or {cost: value} for one variable), cells
    not listed take `default`; a cell listed twice keeps the LAST cost, as in the reference."""
    doms = [domain_values[var_domain[i]] for i in scope]
    shape = tuple(len(d) for d in doms)
    arity = len(scope)
    strides = [int(np.prod(shape[k + 1:])) for k in range(arity)]
    default = c.get("default")
    t = np.full(int(np.prod(shape)), np.nan if default is None else float(default), dtype=np.float64)
    covered = np.zeros(len(t), dtype=bool) if default is None else None
    lut = []
    for i in scope:
        d = var_domain[i]
        if d not in luts:
            luts[d] = _value_lookup(domain_values[d])
        lut.append(luts[d])
    for value, assignments in (c.get("values") or {}).items():
        if arity == 1 and not isinstance(assignments, str):
            try:
                lin = [doms[0].index(assignments)]
            except ValueError:
                raise DcopFormatError(f"{assignments} is not in the domain {var_domain[scope[0]]}") from None
        else:
            toks = [p.split() for p in str(assignments).split("|")]
            for tk in toks:
                if len(tk) != arity:
                    raise DcopFormatError(f"constraint {c_name}: assignment '{' '.join(tk)}' does "
                                          f"not list {arity} values")
            flat = list(itertools.chain.from_iterable(toks))
            lin = np.zeros(len(toks), dtype=np.int64)
            for k in range(arity):
                try:
                    lin += np.fromiter((lut[k][tok] for tok in flat[k::arity]), dtype=np.int64,
                                       count=len(toks)) * strides[k]
                except KeyError as e:
                    raise DcopFormatError(f"{e.args[0]} is not in the domain "
                                          f"{var_domain[scope[k]]}") from None
        t[lin] = value
        if covered is not None:
            covered[lin] = True
    if covered is not None and not covered.all():
        raise DcopFormatError(f"constraint {c_name}: assignments without a cost and no default")
    return t


def _assemble(dom_size, factor_ptr, edge_var, tables, unary, init_value):
    factor_ptr = np.asarray(factor_ptr, dtype=np.int64)
    edge_var = np.asarray(edge_var, dtype=np.int32)
    sizes = np.array([len(t) for t in tables], dtype=np.int64)
    table_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    flat = np.concatenate(tables) if len(tables) else np.zeros(0, dtype=np.float64)
    V = len(dom_size)
    # incident edges of every variable in constraint (= edge id) order, factor_graph.py:277-280
    from .layout import stable_group_order
    order, var_ptr = stable_group_order(edge_var, V)
    return dict(dom_size=np.asarray(dom_size, dtype=np.int32), factor_ptr=factor_ptr,
                edge_var=edge_var, table_off=table_off, tables=flat,
                unary=np.asarray(unary, dtype=np.float64), var_ptr=var_ptr, var_edge=order,
                init_value=np.asarray(init_value, dtype=np.int32))


def load_yaml(filenames: Union[str, os.PathLike, Iterable], seed: Optional[int] = None) -> DcopArrays:
    """One file or several whose contents are concatenated before parsing
    (load_dcop_from_file, yamldcop.py:62-96); `source:` paths are relative to the first file."""
    if isinstance(filenames, (str, os.PathLike)):
        filenames = [filenames]
    text, main_dir = "", None
    for fn in filenames:
        if main_dir is None:
            main_dir = os.path.dirname(os.path.abspath(fn))
        with open(fn, encoding="utf-8") as f:
            text += f.read()
    return loads_yaml(text, main_dir, seed)


# --------------------------------------------------------------------------------------------
# pyDcop objects
# --------------------------------------------------------------------------------------------
def tabulate_constraint(constraint, cache: Optional[TableCache] = None) -> np.ndarray:
    """Flat row-major table of a pyDcop constraint object, axis i <-> constraint.dimensions[i].
    NAryMatrixRelation: its own matrix (relations.py:716-733).  NAryFunctionRelation built from
    an expression (relations.py:1293-1307): vectorised through `tabulate_expression`.  Anything
    else: one call per assignment."""
    dims = list(constraint.dimensions)
    doms = [list(v.domain) for v in dims]
    shape = tuple(len(d) for d in doms)
    m = getattr(constraint, "_m", None)
    if m is not None and tuple(np.shape(m)) == shape:
        return np.asarray(m, dtype=np.float64).reshape(-1)
    names = [v.name for v in dims]
    cache = cache if cache is not None else TableCache()
    f = getattr(constraint, "function", None)
    text = getattr(f, "expression", None)
    if isinstance(text, str) and not getattr(f, "_fixed_vars", None):
        # built from a string (relations.py:1293-1307): re-parse it and skip the per-call
        # argument checks of ExpressionFunction.__call__ (expressionfunction.py:127-144)
        try:
            expr = cache.expression(text, getattr(f, "_source_file", None))
        except (SyntaxError, OSError):
            expr = None
        if expr is not None and sorted(expr.variable_names) == sorted(names):
            perm = tuple(expr.variable_names.index(n) for n in names)
            key = (expr.canonical_key, perm, tuple(tuple(map(repr, d)) for d in doms))
            t = cache.tables.get(key)
            if t is None:
                t = tabulate_expression(expr.func, names, doms, stats=cache.stats).reshape(-1)
                cache.tables[key] = t
            else:
                cache.stats["shared"] += 1
            return t
    return tabulate_expression(lambda **kw: constraint(**kw), names, doms,
                               stats=cache.stats).reshape(-1)


def from_dcop(dcop, variables: Optional[Iterable] = None,
              constraints: Optional[Iterable] = None) -> DcopArrays:
    """Arrays from a pyDcop DCOP (or explicit variable / constraint lists, the signature of
    factor_graph.build_computation_graph, factor_graph.py:245-288).  The per-variable dependent
    constraint search is one stable sort over the edge list instead of |V| scans of |C|."""
    if dcop is not None:
        variables = list(dcop.variables.values())
        constraints = list(dcop.constraints.values())
        name, objective = dcop.name, dcop.objective
    else:
        if variables is None or constraints is None:
            raise ValueError("Constraints AND variables parameters must be provided when not "
                             "building from a dcop")
        variables, constraints = list(variables), list(constraints)
        name, objective = "dcop", "min"
    var_names = [v.name for v in variables]
    vidx = {n: i for i, n in enumerate(var_names)}
    domain_values, var_domain, unary, init_value = {}, [], [], []
    for v in variables:
        dvals = list(v.domain)
        dname = getattr(v.domain, "name", None) or f"d_{v.name}"
        if dname in domain_values and domain_values[dname] != dvals:
            dname = f"{dname}__{v.name}"
        domain_values[dname] = dvals
        var_domain.append(dname)
        cost = getattr(v, "cost_for_val", None)
        unary.extend(float(cost(x)) if cost else 0.0 for x in dvals)
        iv = getattr(v, "initial_value", None)
        init_value.append(dvals.index(iv) if iv is not None else -1)
    cache = TableCache()
    con_names, factor_ptr, edge_var, tables = [], [0], [], []
    for c in constraints:
        for v in c.dimensions:
            if v.name not in vidx:
                raise DcopFormatError(f"constraint {c.name} depends on unknown variable {v.name}")
            edge_var.append(vidx[v.name])
        factor_ptr.append(len(edge_var))
        con_names.append(c.name)
        tables.append(tabulate_constraint(c, cache))
    dom_size = np.array([len(domain_values[d]) for d in var_domain], dtype=np.int32)
    arrays = _assemble(dom_size, factor_ptr, edge_var, tables, unary, init_value)
    return DcopArrays(name=name, objective=objective, arrays=arrays, var_names=var_names,
                      con_names=con_names, domain_values=domain_values, var_domain=var_domain,
                      meta={"tabulation": dict(cache.stats)})


def from_arrays(inst: Dict[str, np.ndarray], name="dcop", objective="min") -> DcopArrays:
    """Wrap generator output (pydcop_b200.generators) so it can be saved / solved by name:
    variables v0.., constraints c0.., domains 0..d-1."""
    a = dict(inst)
    dom = np.asarray(a["dom_size"], dtype=np.int32)
    fp = np.asarray(a["factor_ptr"], dtype=np.int64)
    ev = np.asarray(a["edge_var"], dtype=np.int32)
    if a.get("table_off") is None:
        if len(fp) > 1 and (np.diff(fp) > 0).all():
            ts = np.multiply.reduceat(dom[ev].astype(np.int64), fp[:-1])
        else:
            ts = np.ones(len(fp) - 1, dtype=np.int64)
            np.multiply.at(ts, np.repeat(np.arange(len(fp) - 1), np.diff(fp)), dom[ev].astype(np.int64))
        a["table_off"] = np.concatenate([[0], np.cumsum(ts)]).astype(np.int64)
    if a.get("unary") is None:
        a["unary"] = np.zeros(int(dom.sum()))
    if a.get("var_ptr") is None or a.get("var_edge") is None:
        from .layout import stable_group_order
        a["var_edge"], a["var_ptr"] = stable_group_order(ev, len(dom))
    if a.get("init_value") is None:
        a["init_value"] = np.full(len(dom), -1, dtype=np.int32)
    a["dom_size"], a["factor_ptr"], a["edge_var"] = dom, fp, ev
    sizes = [int(d) for d in np.unique(dom)]
    return DcopArrays(name=name, objective=objective, arrays={k: a[k] for k in ARRAY_KEYS},
                      var_names=GeneratedNames("v", len(dom)), con_names=GeneratedNames("c", len(fp) - 1),
                      domain_values={f"d{d}": list(range(d)) for d in sizes},
                      var_domain=([f"d{sizes[0]}"] * len(dom) if len(sizes) == 1 else [f"d{int(d)}" for d in dom]))


# --------------------------------------------------------------------------------------------
# binary container
# --------------------------------------------------------------------------------------------
MAGIC = b"PDCOPFG1"
_ALIGN = 4096


def save_instance(path: Union[str, os.PathLike], dcop: DcopArrays, table_dtype=None,
                  names: bool = True) -> int:
    """Write `dcop` as: MAGIC, u64 header length, JSON header, then every array raw at a
    4 KiB-aligned offset (so `load_instance` can memory-map it and a reader can `cudaMemcpy`
    straight from the page cache).  Tables are stored as `table_dtype`; the default keeps the source
    dtype (float64 for YAML input: a later f64 solve then sees the costs the YAML solve sees; pass
    float32 to halve the file when every cost is representable).
    `names=False` drops variable / constraint names (they dominate the header at 10^6 scale;
    `load_instance` regenerates v<i> / c<i>).  Returns the file size."""
    arrs = {}
    for k in ARRAY_KEYS:
        a = np.ascontiguousarray(dcop.arrays[k])
        if k == "tables" and table_dtype is not None:
            a = np.ascontiguousarray(a, dtype=table_dtype)
        arrs[k] = a
    header = {"format": 1, "name": dcop.name, "objective": dcop.objective,
              "n_vars": dcop.n_vars, "n_constraints": dcop.n_constraints,
              "domain_values": dcop.domain_values, "meta": dcop.meta, "arrays": {}}
    doms = sorted(dcop.domain_values)
    header["domains"] = doms
    didx = {d: i for i, d in enumerate(doms)}
    arrs["var_domain_id"] = np.array([didx[d] for d in dcop.var_domain], dtype=np.int32)
    if names:
        header["var_names"], header["con_names"] = list(dcop.var_names), list(dcop.con_names)

    def layout(base):
        off = base
        for k, a in arrs.items():
            off = -(-off // _ALIGN) * _ALIGN
            header["arrays"][k] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": off}
            off += a.nbytes
        return off

    # the header holds the offsets, and the offsets depend on the header's length: reserve room
    # for the longest possible offsets (20 digits each) on top of the header written with zeros
    layout(0)
    for d in header["arrays"].values():
        d["offset"] = 0
    hlen = len(json.dumps(header).encode("utf-8")) + 20 * len(arrs) + 64
    end = layout(16 + hlen)
    blob = json.dumps(header).encode("utf-8")
    assert len(blob) <= hlen
    blob = blob.ljust(hlen, b" ")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", hlen))
        f.write(blob)
        for k, a in arrs.items():
            f.seek(header["arrays"][k]["offset"])
            f.write(memoryview(a).cast("B"))
        f.truncate(max(end, f.tell()))
    return end


def load_instance(path: Union[str, os.PathLike], mmap: bool = True) -> DcopArrays:
    """Read a container written by `save_instance`.  With `mmap` the arrays are read-only views
    of the file (no copy until the engine uploads them)."""
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) < 16 or head[:8] != MAGIC:
            raise DcopFormatError(f"{path}: not a pydcop_b200 instance file")
        (hlen,) = struct.unpack("<Q", head[8:])
        try:
            header = json.loads(f.read(hlen).decode("utf-8"))
        except (UnicodeDecodeError, json.JSONDecodeError) as e:
            raise DcopFormatError(f"{path}: corrupt header") from e
    if header.get("format") != 1:
        raise DcopFormatError(f"{path}: unsupported format {header.get('format')!r}")
    size = os.path.getsize(path)
    arrs = {}
    for k, d in header["arrays"].items():
        dt, shape = np.dtype(d["dtype"]), tuple(d["shape"])
        n = int(np.prod(shape)) if shape else 1
        if d["offset"] + n * dt.itemsize > size:
            raise DcopFormatError(f"{path}: truncated (array {k})")
        if n == 0:
            arrs[k] = np.zeros(shape, dtype=dt)
        elif mmap:
            arrs[k] = np.memmap(path, dtype=dt, mode="r", offset=d["offset"], shape=shape)
        else:
            arrs[k] = np.fromfile(path, dtype=dt, count=n, offset=d["offset"]).reshape(shape)
    doms = header["domains"]
    var_domain = [doms[i] for i in arrs.pop("var_domain_id")]
    nv, nc = header["n_vars"], header["n_constraints"]
    return DcopArrays(name=header["name"], objective=header["objective"],
                      arrays={k: arrs[k] for k in ARRAY_KEYS},
                      var_names=header.get("var_names") or GeneratedNames("v", nv),
                      con_names=header.get("con_names") or GeneratedNames("c", nc),
                      domain_values=header["domain_values"], var_domain=var_domain,
                      meta=header.get("meta", {}))


# --------------------------------------------------------------------------------------------
# MaxSum noise (algorithm parameter, not part of the instance)
# --------------------------------------------------------------------------------------------
def stdlib_uniform_stream(n: int, seed: int) -> np.ndarray:
    """The first n values of `random.Random(seed).random()` for an int seed, vectorised: CPython seeds its
    MT19937 with init_by_array over the 32-bit words of abs(seed) and builds each double from two outputs
    (a >> 5, b >> 6) exactly like numpy's legacy RandomState.random_sample, so the two streams are
    bit-identical (checked against the stdlib in tests/test_ingest.py)."""
    a, words = abs(int(seed)), []
    while True:
        words.append(a & 0xFFFFFFFF)
        a >>= 32
        if not a:
            break
    return np.random.RandomState(words).random_sample(int(n))     # a LIST: a 1-element array would be read as a scalar


def add_noise(unary: np.ndarray, noise: float, seed: Optional[int] = None) -> np.ndarray:
    """MaxSum wraps every variable in VariableNoisyCostFunc(noise_level=noise) unless noise == 0
    (maxsum.py:474-483): cost + U(0, noise) per (variable, value), drawn in variable order then
    value order from Python's `random` (objects.py:566-567): uniform(0, noise) = 0 + noise * random().
    `seed` makes the draws repeatable — and equal to what `random.Random(seed)` would draw one by one."""
    if not noise:
        return np.asarray(unary, dtype=np.float64)
    n = len(unary)
    if seed is not None:
        draws = 0.0 + (float(noise) - 0.0) * stdlib_uniform_stream(n, seed)
    else:
        draws = np.array([random.uniform(0, noise) for _ in range(n)])
    return np.asarray(unary, dtype=np.float64) + draws
