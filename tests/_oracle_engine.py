"""Test double with the MaxSumEngine / DsaEngine driving API (init / step / values), backed by the
CPU oracle.  Lets the host logic around the engines run on a box without a GPU.  Test-only."""
import numpy as np

import oracle as orc


class OracleEngine:
    def __init__(self, kind, layout, inst, params):
        self.kind = kind
        if kind == "maxsum":
            keys = ("damping", "damping_nodes", "stability", "start_messages")
            self.o = orc.MaxSumOracle(inst, np.float64, mode=params["mode"],
                                      **{k: params[k] for k in keys if k in params})
        elif kind == "mgm":
            keys = ("stop_cycle", "seed", "break_mode")
            self.o = orc.MgmOracle(dict(inst), np.float64, mode=params["mode"],
                                   **{k: params[k] for k in keys if k in params})
        else:
            keys = ("probability", "p_mode", "variant", "stop_cycle", "seed")
            self.o = orc.DsaOracle(dict(inst), np.float64, mode=params["mode"], var_costs=(kind == "adsa"),
                                   **{k: params[k] for k in keys if k in params})

    def init(self):
        self.o.init()
        return self

    def step(self, n=1):
        self.o.step(n)
        return self

    @property
    def finished(self):
        return bool(getattr(self.o, "finished", False))

    def values(self):
        if self.kind == "mgm":
            cost = self.o.cost.astype(np.float64)
            cost[self.o.has_cost == 0] = np.nan
            return self.o.val.copy(), cost
        if self.kind == "maxsum":
            return self.o.value.copy(), self.o.value_cost.copy()
        return self.o.val.copy()
