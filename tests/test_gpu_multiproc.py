"""GPU (>= 2 devices): the real multi-process path — one process per GPU, NCCL for the plumbing — in
both halo modes: pack -> all_to_all -> unpack, and the NVLink peer stores through CUDA IPC (FUSED into the
factor-side / variable-side kernels by default; the separate push kernels behind PYDCOP_B200_PUSH_FUSED=0) closed by
the DEVICE-SIDE epoch barrier (csrc/peer_sync.cuh, fg_maxsum_shard_step / fg_dsa_shard_step).  The
all-gathered assignment must equal the single-GPU engine's, which the other tests tie to the oracle
bit for bit.  Cases cover: several step() calls and a re-init on one engine (the epoch counter only
grows), the joined push (one launch after both sides instead of r rows behind the factor side / q rows behind
the variable side), the per-row push (no destination runs), and a
deliberately IMBALANCED partition — a fast rank's push must not be overwritten by a slow rank."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _owner(case, n_vars, world):
    """Partition argument of a case: a method name, or 'imbalanced' = rank 0 owns 85 % of the ids."""
    part = case.get("partition", "blocks")
    if part != "imbalanced":
        return part
    own = np.zeros(n_vars, dtype=np.int32)
    tail = n_vars - int(n_vars * 0.85)
    own[n_vars - tail:] = 1 + (np.arange(tail) % (world - 1))
    return own


def _maxsum_worker(rank, world, port, case, q):
    try:
        for k, v in case.get("env", {}).items():
            os.environ[k] = v
        import torch
        import torch.distributed as dist
        from pydcop_b200 import MaxSumEngine, build_layout
        from pydcop_b200.generators import random_factor_graph
        from pydcop_b200.multigpu import ShardedMaxSum
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        n = case.get("n_vars", 2000)
        inst = random_factor_graph(n, 10, 2 * n, 2, seed=21)
        sh = ShardedMaxSum(inst, rank, world, dev, precision=case.get("precision", "f32"), halo=case["mode"],
                           partition=_owner(case, n, world)).init()
        plan = case.get("steps", [9])
        ok, got = True, None
        for i, steps in enumerate(plan):
            if steps == "init":
                sh.init()
                continue
            sh.step(steps)
        sh.check()
        got = sh.values()
        cost = sh.solution_cost(9.0, inst["unary"])     # table entries equal to 9 count as violations
        used = "p2p" if sh.peer is not None else "nccl"
        if sh.peer is not None and case.get("partition", "blocks") != "imbalanced":
            # default: push kernels behind each side; PYDCOP_B200_PUSH_FUSED=1: the warp
            # kernels store the boundary rows themselves (all classes binary d=10, degrees <= 16)
            assert sh.peer.fused == case.get("fused", False), sh.peer.fused
        if rank == 0:
            cur = 0
            for s in plan:     # cycles since the last init
                cur = 0 if s == "init" else cur + s
            ref = MaxSumEngine(build_layout(**inst), device=dev, precision=case.get("precision", "f32")).init().step(cur)
            ok = bool(np.array_equal(got, ref.values()[0]))
            want = ref.solution_cost(9.0, inst["unary"])
            ok = ok and cost[1] == want[1] and cost[1] > 0 and abs(cost[0] - want[0]) <= 1e-6 * abs(want[0])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok" if ok else "MISMATCH", used))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc(), "?"))


def _dsa_worker(rank, world, port, case, q):
    try:
        import torch
        import torch.distributed as dist
        from pydcop_b200.engine import DsaEngine
        from pydcop_b200.generators import random_factor_graph
        from pydcop_b200.layout import build_layout
        from pydcop_b200.multigpu_dsa import ShardedDsa
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        n = case.get("n_vars", 4000)
        inst = random_factor_graph(n, 20, 3 * n, 2, seed=22, noise=0.0)
        inst["tables"] = np.floor(inst["tables"] / 3.0).astype(np.float32)
        kw = dict(precision="f32", variant="B", seed=9, stop_cycle=case.get("stop_cycle", 0))
        sh = ShardedDsa(inst, rank, world, dev, halo=case["mode"], partition=_owner(case, n, world), **kw).init()
        for steps in case.get("steps", [10]):
            sh.step(steps)
        sh.check()
        got = sh.values()
        cost = sh.solution_cost(2.0, inst["unary"])     # table entries equal to 2 count as violations
        used = "p2p" if sh.peer is not None else "nccl"
        ok = True
        if rank == 0:
            ref = DsaEngine(build_layout(**inst), device=dev, **kw).init().step(sum(case.get("steps", [10])))
            ok = bool(np.array_equal(got, ref.values()))
            want = ref.solution_cost(2.0, inst["unary"])
            ok = ok and cost[1] == want[1] and cost[1] > 0 and abs(cost[0] - want[0]) <= 1e-6 * max(1.0, abs(want[0]))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok" if ok else "MISMATCH", used))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc(), "?"))


def _run(worker, world, case):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = [q.get(timeout=150) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    assert all(r[1] == "ok" for r in results), results
    assert all(r[2] == case["mode"] for r in results), results


MAXSUM_CASES = {
    "nccl": dict(mode="nccl"),
    "p2p": dict(mode="p2p"),
    "p2p-steps-reinit": dict(mode="p2p", steps=[1, 3, "init", 2, 5, 4]),
    "p2p-fused": dict(mode="p2p", env={"PYDCOP_B200_PUSH_FUSED": "1"}, steps=[4, 5], fused=True),
    "p2p-early-q-push": dict(mode="p2p", env={"PYDCOP_B200_PUSH_EARLY": "1"}, steps=[4, 5]),
    "p2p-early-q-push-third-stream": dict(mode="p2p", env={"PYDCOP_B200_PUSH_EARLY": "2"}, steps=[4, 5]),
    "p2p-unchained": dict(mode="p2p", env={"PYDCOP_B200_PUSH_CHAIN": "0"}, steps=[4, 5]),
    "p2p-joined-push": dict(mode="p2p", env={"PYDCOP_B200_PUSH_SPLIT": "0"}, steps=[4, 5]),
    "p2p-per-row-push": dict(mode="p2p", env={"PYDCOP_B200_PUSH_RUNS": "0"}, steps=[4, 5]),
    "p2p-multilevel-f64": dict(mode="p2p", partition="multilevel", precision="f64"),
    "p2p-imbalanced": dict(mode="p2p", partition="imbalanced", n_vars=20000, steps=[12]),
}


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", sorted(MAXSUM_CASES))
def test_sharded_maxsum_processes_match_single_gpu(name, world):
    _run(_maxsum_worker, world, MAXSUM_CASES[name])


DSA_CASES = {
    "nccl": dict(mode="nccl", partition="multilevel"),
    "p2p": dict(mode="p2p", partition="multilevel"),
    "p2p-steps": dict(mode="p2p", steps=[1, 2, 7]),
    "p2p-stop-cycle": dict(mode="p2p", stop_cycle=6, steps=[4, 4, 2]),
    # rank 0 owns 85 % of the variables: the other ranks finish their kernel and push long before rank 0's
    # kernel reaches its ghost entries (a local write of the ghost would overwrite the push)
    "p2p-imbalanced": dict(mode="p2p", partition="imbalanced", n_vars=200000, steps=[12]),
}


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", sorted(DSA_CASES))
def test_sharded_dsa_processes_match_single_gpu(name, world):
    _run(_dsa_worker, world, DSA_CASES[name])
