"""Peer memory plumbing shared by the multi-GPU MaxSum and DSA paths: CUDA-IPC mapping of the other
ranks' buffers into this process and the device-side epoch barrier (include/pydcop_b200.h,
fg_peer_sync_t / fg_halo_plan_t; kernels in csrc/peer_sync.cuh).

One rank = one process = one GPU.  `PeerMap.map(tensors)` returns, for every rank, the address at
which that rank's tensors are reachable from THIS process (own tensors: their own addresses).
`PeerSync` owns the epoch flags: an int64 array with one slot per rank in this rank's memory; slot p
is written only by rank p (st.release.sys over NVLink) and read only by this rank
(ld.acquire.sys).  It replaces the per-cycle NCCL all_reduce of round 1: no host round trip, no
collective — the cycle_id handshake of SynchronousComputationMixin
(pydcop/infrastructure/computations.py:696-718) restated for GPUs that share an NVSwitch.
"""
import ctypes as C

import numpy as np

from . import _cabi


class PeerMap:
    """CUDA IPC handles exchanged through torch.distributed (object all-gather, any backend)."""

    def __init__(self, lib, device, rank, world, group=None):
        self.lib, self.device, self.rank, self.world, self.group = lib, device, rank, world, group
        self._mapped = {}     # (rank, handle bytes) -> base address in this process

    def map(self, tensors):
        """int64 array [world, len(tensors)]: address of every rank's i-th tensor in this process."""
        import torch
        import torch.distributed as dist
        lib = self.lib
        mine = []
        for t in tensors:
            hb = (C.c_ubyte * 64)()
            off = C.c_int64()
            rc = lib.fg_ipc_export(C.c_void_p(t.data_ptr()), C.cast(hb, C.c_void_p), C.byref(off))
            if rc != 0:
                raise RuntimeError(f"fg_ipc_export failed rc={rc}")
            mine.append((bytes(hb), int(off.value)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        base = np.zeros((self.world, len(tensors)), dtype=np.int64)
        with torch.cuda.device(self.device):
            for rnk in range(self.world):
                if rnk == self.rank:
                    base[rnk] = [t.data_ptr() for t in tensors]
                    continue
                for i, (hbytes, off) in enumerate(everyone[rnk]):
                    key = (rnk, hbytes)
                    if key not in self._mapped:
                        out = C.c_void_p()
                        buf = (C.c_ubyte * 64).from_buffer_copy(hbytes)
                        rc = lib.fg_ipc_import(C.cast(buf, C.c_void_p), C.byref(out))
                        if rc != 0:
                            raise RuntimeError(f"fg_ipc_import failed rc={rc} (rank {rnk})")
                        self._mapped[key] = int(out.value)
                    base[rnk, i] = self._mapped[key] + off
        return base


def peer_slot_addresses(flag_base, my_rank, peers):
    """Address of slot `my_rank` inside each peer's flag array (8-byte slots).  flag_base[p] = address
    of rank p's flag array as mapped into this process.  Pure function (tested on CPU)."""
    flag_base = np.asarray(flag_base, dtype=np.int64)
    return [int(flag_base[p]) + 8 * int(my_rank) for p in peers]


class PeerSync:
    """Epoch flags of one rank + the struct the C side takes.  `peers`: the ranks this rank exchanges
    boundary data with (it releases to exactly those and waits for exactly those)."""

    def __init__(self, pmap: PeerMap, peers, timeout_s=20.0):
        import torch
        import torch.distributed as dist
        self.pmap = pmap
        dev = pmap.device
        peers = [int(p) for p in peers if int(p) != pmap.rank]
        if len(peers) > _cabi.FG_MAX_PEERS:
            raise RuntimeError(f"{len(peers)} peers > FG_MAX_PEERS")
        self.peers = peers
        self.flags = torch.zeros(max(pmap.world, 1), dtype=torch.int64, device=dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        base = pmap.map([self.flags])[:, 0]
        s = _cabi.FgPeerSync()
        s.n_peers, s.my_rank = len(peers), pmap.rank
        s.dev_flags = self.flags.data_ptr()
        for i, (p, a) in enumerate(zip(peers, peer_slot_addresses(base, pmap.rank, peers))):
            s.peer_rank[i] = p
            s.peer_slot[i] = a
        s.dev_error = self.error.data_ptr()
        s.timeout_ns = int(timeout_s * 1e9)
        self.struct = s
        dist.barrier(group=pmap.group)   # every rank's flags exist and are zero before anyone releases

    def check(self):
        """Raise if a device-side wait timed out (reads one int32 from the device)."""
        e = int(self.error.item())
        if e:
            raise _cabi.EngineError(f"device-side cycle barrier: no signal from rank {e - 1} within the timeout")
