"""Direct solve: a DCOP file (or arrays, or a pyDcop DCOP object) -> GPU engine -> result dict
(algorithms: maxsum, dsa, adsa, mgm).

The plugin modules (pydcop_b200/algorithms/) keep pyDcop's orchestrator, agents and one Python
computation object per graph node around the engine — that is the drop-in path, and it is what
bounds instance size (10^6 proxy objects and agent threads do not fit).  This entry skips the
control plane for large instances: ingestion (pydcop_b200.ingest) -> layout -> engine, and returns
the same result keys `pydcop solve` prints (pydcop/infrastructure/orchestrator.py:1262-1272,
pydcop/commands/solve.py:611-624): status, assignment, cost, violation, time, cycle, msg_count,
msg_size.

Parameters and defaults are the reference algorithms' own (maxsum.py:212-220, dsa.py:130-135,
mgm.py:78-81).
MaxSum has no termination test of its own (maxsum.py: it runs until `stop_cycle` or the
orchestrator's timeout), so one of `stop_cycle` / `timeout` is required here as well; status is
FINISHED when `stop_cycle` was reached and TIMEOUT otherwise, as in the reference
(orchestrator.py:1262, commands/solve.py:560-580).

No CPU path: the engines raise EngineError without a CUDA device.  `engine_factory` exists so
the host logic can be tested without one.
"""
import json
import os
import time
from typing import Any, Callable, Dict, Optional

import numpy as np

from . import ingest
from .layout import build_layout

MAXSUM_DEFAULTS = {"damping": 0.5, "damping_nodes": "both", "stability": 0.1, "noise": 0.01,
                   "start_messages": "leafs", "stop_cycle": 0}
DSA_DEFAULTS = {"probability": 0.7, "p_mode": "fixed", "variant": "B", "stop_cycle": 0}
ADSA_DEFAULTS = {"period": 0.5, "probability": 0.7, "variant": "B", "stop_cycle": 0}   # adsa.py:121-125 + stop_cycle
MGM_DEFAULTS = {"break_mode": "lexic", "stop_cycle": 0}
DEFAULTS = {"maxsum": MAXSUM_DEFAULTS, "dsa": DSA_DEFAULTS, "mgm": MGM_DEFAULTS, "adsa": ADSA_DEFAULTS}
_CHOICES = {"damping_nodes": ("vars", "factors", "both", "none"),
            "start_messages": ("leafs", "leafs_vars", "all"),
            "p_mode": ("fixed", "arity"), "variant": ("A", "B", "C"),
            "break_mode": ("lexic", "random")}
ALGOS = {"maxsum": "maxsum", "maxsum_gpu": "maxsum", "dsa": "dsa", "dsa_gpu": "dsa",
         "mgm": "mgm", "mgm_gpu": "mgm", "adsa": "adsa", "adsa_gpu": "adsa"}


def check_params(kind: str, params: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    """Defaults + type / choice validation, the job of AlgoParameterDef
    (pydcop/algorithms/__init__.py:180-290): unknown names and invalid values raise ValueError."""
    defaults = DEFAULTS[kind]
    out = dict(defaults)
    for k, v in (params or {}).items():
        if k not in defaults:
            raise ValueError(f"Unknown parameter for algorithm {kind}: {k}")
        want = type(defaults[k])
        try:
            v = want(v) if want is not str else str(v)
        except (TypeError, ValueError):
            raise ValueError(f"Invalid value for parameter {k}: {v!r}") from None
        if k in _CHOICES and v not in _CHOICES[k]:
            raise ValueError(f"Invalid value for parameter {k}: {v!r} not in {_CHOICES[k]}")
        out[k] = v
    return out


def load(problem, seed: Optional[int] = None) -> ingest.DcopArrays:
    """DcopArrays from: a DcopArrays, a path / list of paths (YAML, or the binary container by its
    magic), a dict of front-door arrays, or a pyDcop DCOP object."""
    if isinstance(problem, ingest.DcopArrays):
        return problem
    if isinstance(problem, dict):
        return ingest.from_arrays(problem)
    if isinstance(problem, (str, os.PathLike)):
        with open(problem, "rb") as f:
            magic = f.read(len(ingest.MAGIC))
        if magic == ingest.MAGIC:
            return ingest.load_instance(problem)
        return ingest.load_yaml(problem, seed)
    if isinstance(problem, (list, tuple)):
        return ingest.load_yaml(problem, seed)
    if hasattr(problem, "variables") and hasattr(problem, "constraints"):
        return ingest.from_dcop(problem)
    raise TypeError(f"cannot load a DCOP from {type(problem).__name__}")


def isolated_values(dcop: ingest.DcopArrays, mode: str) -> np.ndarray:
    """DSA start value of a variable without NEIGHBOURS — no constraint at all, or only unary ones (the engines
    treat both as isolated and take this value verbatim): argopt of (own cost, value) in Python tuple order over
    the real domain values (dsa.py:278-289, relations.py:1641-1669)."""
    a = dcop.arrays
    out = np.zeros(dcop.n_vars, dtype=np.int32)
    uoff = np.concatenate([[0], np.cumsum(a["dom_size"].astype(np.int64))])
    fp, ev = np.asarray(a["factor_ptr"], dtype=np.int64), np.asarray(a["edge_var"], dtype=np.int64)
    n_nbr = np.zeros(dcop.n_vars, dtype=np.int64)
    if len(ev):
        np.add.at(n_nbr, ev, np.repeat(np.diff(fp) - 1, np.diff(fp)))
    for i in np.nonzero(n_nbr == 0)[0]:
        dom = dcop.values_of(int(i))
        pairs = [(float(a["unary"][uoff[i] + k]), x) for k, x in enumerate(dom)]
        try:
            best = min(pairs) if mode == "min" else max(pairs)
            out[i] = dom.index(best[1])
        except TypeError:  # unorderable values: first optimum
            cs = [p[0] for p in pairs]
            out[i] = int(np.argmin(cs) if mode == "min" else np.argmax(cs))
    return out


def solution_cost(dcop: ingest.DcopArrays, value_index, infinity: float = float("inf")):
    """(violation, cost): constraints and variable costs equal to `infinity` are counted, the
    others summed (pydcop/dcop/dcop.py:319-367; `infinity` default as commands/solve.py `-i`: float("inf"))."""
    a = dcop.arrays
    idx = np.asarray(value_index, dtype=np.int64)
    fp, ev = a["factor_ptr"].astype(np.int64), a["edge_var"].astype(np.int64)
    dom = a["dom_size"].astype(np.int64)
    lin = np.zeros(len(fp) - 1, dtype=np.int64)
    arity = np.diff(fp)
    for j in range(int(arity.max(initial=0))):
        m = arity > j
        e = fp[:-1][m] + j
        lin[m] = lin[m] * dom[ev[e]] + idx[ev[e]]
    costs = np.concatenate([
        np.asarray(a["tables"])[np.asarray(a["table_off"][:-1], dtype=np.int64) + lin].astype(np.float64),
        np.asarray(a["unary"], dtype=np.float64)[np.concatenate([[0], np.cumsum(dom)])[:-1] + idx]])
    hard = costs == infinity
    return int(hard.sum()), float(costs[~hard].sum())


def name_rank(dcop: ingest.DcopArrays) -> np.ndarray:
    """Position of every variable's name in sorted order: MGM's tie break (mgm.py:574-583)."""
    order = sorted(range(dcop.n_vars), key=lambda i: dcop.var_names[i])
    rank = np.zeros(dcop.n_vars, dtype=np.int32)
    rank[order] = np.arange(dcop.n_vars, dtype=np.int32)
    return rank


def _default_engine(kind, layout, dcop, params, mode, precision, device, seed):
    from .engine import DsaEngine, MaxSumEngine, MgmEngine
    if kind == "mgm":
        return MgmEngine(layout, device=device, precision=precision, mode=mode,
                         stop_cycle=params["stop_cycle"], seed=seed or 0,
                         break_mode=params["break_mode"], var_rank=name_rank(dcop),
                         isolated_value=isolated_values(dcop, mode))
    if kind == "maxsum":
        return MaxSumEngine(layout, device=device, precision=precision, mode=mode,
                            damping=params["damping"], damping_nodes=params["damping_nodes"],
                            stability=params["stability"], start_messages=params["start_messages"],
                            record_sent=False)
    return DsaEngine(layout, device=device, precision=precision, mode=mode,
                     probability=params["probability"], p_mode=params.get("p_mode", "fixed"),
                     variant=params["variant"], stop_cycle=params["stop_cycle"], seed=seed or 0,
                     isolated_value=isolated_values(dcop, mode), var_costs=(kind == "adsa"))


def _sharded_engine(kind, inst, dcop, params, mode, precision, device, seed, partition, halo, sharded_kwargs):
    """One rank's engine of a multi-process run (torch.distributed is initialised): the variables
    are split over the ranks (rank 0 computes the partition and broadcasts it), boundary rows /
    values are exchanged every cycle (pydcop_b200.multigpu, multigpu_dsa)."""
    import torch
    import torch.distributed as dist
    from .multigpu import ShardedMaxSum, broadcast_owner
    from .multigpu_dsa import ShardedDsa
    rank, world = dist.get_rank(), dist.get_world_size()
    if kind == "mgm":
        raise ValueError("mgm is not sharded: run it on one GPU")
    kw = dict(sharded_kwargs or {})
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if torch.cuda.is_available() and "engine_factory" not in kw
                                             else torch.device("cpu"))
    if isinstance(partition, str):
        owner, err = broadcast_owner(inst, world, rank, dev, partition)
    else:                                  # an owner array every rank already holds
        owner, err = np.asarray(partition, dtype=np.int32), None
    if kind == "maxsum":
        eng = ShardedMaxSum(inst, rank, world, dev, precision=precision, halo=halo, partition=owner, mode=mode,
                            damping=params["damping"], damping_nodes=params["damping_nodes"],
                            stability=params["stability"], start_messages=params["start_messages"],
                            **dict({"record_sent": False} if "engine_factory" not in kw else {}, **kw))
    else:
        eng = ShardedDsa(inst, rank, world, dev, precision=precision, halo=halo, partition=owner, mode=mode,
                         probability=params["probability"], p_mode=params.get("p_mode", "fixed"),
                         variant=params["variant"], stop_cycle=params["stop_cycle"], seed=seed or 0,
                         isolated_value=isolated_values(dcop, mode), var_costs=(kind == "adsa"), **kw)
    eng.partition_error = err
    return eng


def solve(problem, algo: str = "maxsum", algo_params: Optional[Dict[str, Any]] = None,
          timeout: Optional[float] = None, precision: str = "f32", device=None,
          seed: Optional[int] = None, infinity: float = float("inf"), chunk: int = 50,
          on_cycle: Optional[Callable[[int, np.ndarray], None]] = None,
          engine_factory: Optional[Callable] = None, distributed: Optional[bool] = None,
          partition="auto", halo: str = "auto", sharded_kwargs: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """Run `algo` on `problem` and return pyDcop's result dict.

    Under `torchrun` (torch.distributed initialised, world size > 1; `distributed` None = detect)
    every rank calls solve() with the same arguments: the variables are partitioned over the
    ranks (`partition`: auto | blocks | multilevel | an owner array), each rank runs its shard on
    its GPU and boundary messages / values are exchanged every cycle (`halo`: auto | p2p | nccl);
    every rank returns the same result.  `seed` defaults to 0 there (all ranks must draw the same
    noise).

    problem      see `load`
    algo         maxsum | dsa | mgm (the *_gpu spellings of the plugin modules are accepted)
    algo_params  the reference's parameter names; `stop_cycle` bounds the run in cycles
    timeout      wall-clock bound in seconds (checked every `chunk` cycles)
    precision    f32 | f64 (f64 reproduces the reference's double arithmetic bit for bit)
    seed         MaxSum: the U(0, noise) draws; DSA: the Philox stream of the random choices
    on_cycle     callback(cycle, value_index) after every chunk (metrics collection)
    """
    if algo not in ALGOS:
        raise ValueError(f"unknown algorithm {algo!r}: expected one of {sorted(ALGOS)}")
    kind = ALGOS[algo]
    params = check_params(kind, algo_params)
    stop_cycle = int(params["stop_cycle"])
    if not stop_cycle and timeout is None:
        raise ValueError(f"{kind} does not stop by itself: give algo_params['stop_cycle'] or a timeout")
    if distributed is None:
        try:
            import torch.distributed as dist
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except ImportError:
            distributed = False
    if distributed and seed is None:
        seed = 0
    t0 = time.perf_counter()
    dcop = load(problem, seed)
    mode = dcop.objective
    inst = dcop.instance()
    if kind == "maxsum":
        inst["unary"] = ingest.add_noise(inst["unary"], params["noise"], seed)
    n_gpus = 1
    if distributed:
        import torch.distributed as dist
        engine = _sharded_engine(kind, inst, dcop, params, mode, precision, device, seed, partition, halo,
                                 sharded_kwargs)
        n_gpus = dist.get_world_size()
    else:
        layout = build_layout(**inst)
        if engine_factory is not None:
            engine = engine_factory(kind, layout, dict(inst, var_rank=name_rank(dcop)),
                                    dict(params, mode=mode, seed=seed or 0))
        else:
            engine = _default_engine(kind, layout, dcop, params, mode, precision, device, seed)
    t_packed = time.perf_counter()
    engine.init()
    cycle, status = 0, "FINISHED"
    deadline = None if timeout is None else t0 + float(timeout)
    while True:
        n = chunk if not stop_cycle else min(chunk, stop_cycle - cycle)
        if n <= 0:
            break
        expired = deadline is not None and time.perf_counter() >= deadline
        if distributed and deadline is not None:   # every rank must leave the loop in the same round
            import torch
            flag = torch.tensor([int(expired)], dtype=torch.int32, device=engine.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            expired = bool(flag.item())
        if expired:
            status = "TIMEOUT"
            break
        engine.step(n)
        cycle += n
        if on_cycle is not None:
            on_cycle(cycle, _indices(engine))
    idx = _indices(engine)
    if hasattr(getattr(engine, "engine", engine), "_solution_cost"):   # reduced on the device(s); sharded engines all-reduce over NCCL
        cost, violation = engine.solution_cost(infinity, dcop.arrays["unary"])
    else:                                  # engine seam of the CPU tests
        violation, cost = solution_cost(dcop, idx, infinity)
    return {"status": status, "assignment": dcop.assignment(idx), "cost": cost,
            "violation": violation, "time": time.perf_counter() - t0, "cycle": cycle,
            "msg_count": 0, "msg_size": 0,  # nothing crosses an agent boundary
            "algo": kind, "precision": precision, "n_gpus": n_gpus,
            "ingest_time": t_packed - t0,
            # bytes this process moved host -> device (problem arrays) and device -> host (assignment, costs)
            "h2d_bytes": int(getattr(getattr(engine, "engine", engine), "h2d_bytes", 0)),
            "d2h_bytes": int(len(idx) * 12 + 16)}


def _indices(engine) -> np.ndarray:
    out = engine.values()
    return np.asarray(out[0] if isinstance(out, tuple) else out)


def main(argv=None):
    """`python -m pydcop_b200.solve --algo maxsum -p stop_cycle:100 problem.yaml` — argument names
    follow `pydcop solve` (commands/solve.py:380-470): -a/--algo, -p/--algo_params name:value,
    -t/--timeout; prints the result as JSON."""
    import argparse
    ap = argparse.ArgumentParser(prog="pydcop_b200.solve")
    ap.add_argument("dcop_files", nargs="+")
    ap.add_argument("-a", "--algo", default="maxsum", choices=sorted(ALGOS))
    ap.add_argument("-p", "--algo_params", action="append", default=[], metavar="name:value")
    ap.add_argument("-t", "--timeout", type=float, default=None)
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("-i", "--infinity", type=float, default=float("inf"),
                    help="cost that marks a violated hard constraint; like `pydcop solve -i` it defaults to inf "
                         "(commands/solve.py:315-324: the help text there says 10 000, the default is float('inf'))")
    ap.add_argument("--no-assignment", action="store_true",
                    help="omit the assignment from the output (10^6 variables)")
    ap.add_argument("--save", metavar="FILE", help="also write the instance as a binary container")
    ap.add_argument("--run_metrics", metavar="FILE",
                    help="CSV of cycle,time,cost,violation,msg_count,msg_size,status every --metrics_every "
                         "cycles (the columns of `pydcop solve --run_metrics`, commands/solve.py:356-375)")
    ap.add_argument("--metrics_every", type=int, default=10, metavar="CYCLES")
    ap.add_argument("--partition", default="auto", choices=["auto", "blocks", "multilevel"],
                    help="under torchrun: how the variables are split over the GPUs")
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:   # launched by torchrun: one process per GPU over NCCL
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    params = {}
    for p in args.algo_params:
        if ":" not in p:
            ap.error(f"algo_params must be name:value, got {p!r}")
        k, v = p.split(":", 1)
        params[k] = v
    files = args.dcop_files if len(args.dcop_files) > 1 else args.dcop_files[0]
    dcop = load(files, args.seed)
    if args.save and rank == 0:
        ingest.save_instance(args.save, dcop)
    rows, t_start = [], time.perf_counter()

    def collect(cycle, idx):
        violation, cost = solution_cost(dcop, idx, args.infinity)
        rows.append((cycle, time.perf_counter() - t_start, cost, violation, 0, 0, "RUNNING"))

    res = solve(dcop, args.algo, params, args.timeout, args.precision, seed=args.seed,
                infinity=args.infinity, partition=args.partition,
                chunk=max(1, args.metrics_every) if args.run_metrics else 50,
                on_cycle=collect if args.run_metrics else None)
    if args.run_metrics and rank == 0:
        import csv
        with open(args.run_metrics, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["cycle", "time", "cost", "violation", "msg_count", "msg_size", "status"])
            w.writerows(rows)
            w.writerow([res["cycle"], res["time"], res["cost"], res["violation"], 0, 0, res["status"]])
    if args.no_assignment:
        res.pop("assignment")
    if rank == 0:
        print(json.dumps(res, indent=2, sort_keys=True, default=str))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
