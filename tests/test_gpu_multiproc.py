"""GPU (>= 2 devices): the real multi-process path — one process per GPU, NCCL — in both halo modes
(pack -> all_to_all -> unpack, and the direct NVLink peer push through CUDA IPC).  The all-gathered
assignment must equal the single-GPU engine's, which the other tests tie to the oracle bit-exactly."""
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


def _worker(rank, world, port, mode, q):
    try:
        import torch
        import torch.distributed as dist
        from pydcop_b200 import MaxSumEngine, build_layout
        from pydcop_b200.generators import random_factor_graph
        from pydcop_b200.multigpu import ShardedMaxSum
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        inst = random_factor_graph(2000, 10, 4000, 2, seed=21)
        sh = ShardedMaxSum(inst, rank, world, dev, precision="f32", halo=mode).init().step(9)
        got = sh.values()
        used = "p2p" if sh.peer is not None else "nccl"
        ok = True
        if rank == 0:
            ref = MaxSumEngine(build_layout(**inst), device=dev, precision="f32").init().step(9)
            ok = bool(np.array_equal(got, ref.values()[0]))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok" if ok else "MISMATCH", used))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc(), "?"))


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_two_process_sharded_matches_single_gpu(mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=20)
        if p.is_alive():
            p.kill()
    assert all(r[1] == "ok" for r in results), results
    assert all(r[2] == mode for r in results), results


def _dsa_worker(rank, world, port, mode, q):
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
        inst = random_factor_graph(4000, 20, 12000, 2, seed=22, noise=0.0)
        inst["tables"] = np.floor(inst["tables"] / 3.0).astype(np.float32)
        sh = ShardedDsa(inst, rank, world, dev, precision="f32", variant="B", seed=9, halo=mode,
                        partition="multilevel").init().step(10)
        got = sh.values()
        used = "p2p" if sh.peer is not None else "nccl"
        ok = True
        if rank == 0:
            ref = DsaEngine(build_layout(**inst), device=dev, precision="f32", variant="B", seed=9).init().step(10)
            ok = bool(np.array_equal(got, ref.values()))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok" if ok else "MISMATCH", used))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc(), "?"))


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_two_process_sharded_dsa_matches_single_gpu(mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dsa_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=20)
        if p.is_alive():
            p.kill()
    assert all(r[1] == "ok" for r in results), results
    assert all(r[2] == mode for r in results), results
