#!/usr/bin/env python
"""Extra MaxSum trajectories from the UNMODIFIED reference for cases the main fixture set does not
hold: infinite (hard-constraint) costs.  Same lock-step method as oracle/make_golden.py.

TEST INFRASTRUCTURE (build container only).  Fixtures are written with the prefix `msx_`: the CPU
tests (oracle, kernel source through the host shim) use them; the GPU parity tests, which
enumerate `ms_*`, will pick them up once they have been run on a device for the first time.

    python oracle/make_golden_extra.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    for name, mode, sign in (("msx_hard_inf_min", "min", 1.0), ("msx_hard_inf_max", "max", -1.0)):
        if not want(name):
            continue
        rng = np.random.default_rng(50)
        vs, cs = G.random_instance(rng, 14, [3, 4], 18, [2, 2, 3, 1])
        for k, c in enumerate(cs):       # a third of the constraints forbid some assignments
            if k % 3 == 0:
                m = c._m
                mask = rng.random(m.shape) < 0.3
                m[mask] = sign * np.inf
        arr, meta = G.run_maxsum(vs, cs, {"noise": 0.0}, mode, 20, seed=51)
        G.save(name, arr, meta)


if __name__ == "__main__":
    main()
