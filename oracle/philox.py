"""Philox4x32-10 counter-based RNG, pure Python.  TEST INFRASTRUCTURE / shared draw definition.

The reference's DSA consumes stdlib/numpy RNG streams in thread-arrival order
(pydcop/algorithms/dsa.py:411-412, pydcop/infrastructure/computations.py:1080-1087), which is
not reproducible even CPU-vs-CPU.  Parity is therefore defined on *injected* draws: every
random decision of variable `v` at cycle `c` is a pure function of (seed, v, c).  The same
function is implemented here (patched into the reference by oracle/make_golden.py), in
oracle/dcop_oracle.c and in the CUDA kernel (pydcop_b200/csrc/philox.cuh).

  block = philox4x32_10(counter=(v, c, 0, 0), key=(seed & 0xffffffff, seed >> 32))
  u      = ((block[0] >> 5) * 2**26 + (block[1] >> 6)) / 2**53      # in [0, 1), 53 bits
  choice = (block[2] * n) >> 32                                      # uniform index in [0, n)
The initial random value of DSA uses c = INIT_CYCLE (0xffffffff).
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF
INIT_CYCLE = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = [int(x) & MASK for x in counter]
    k0, k1 = [int(x) & MASK for x in key]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw(seed: int, var: int, cycle: int):
    """Returns (u in [0,1) as float64, raw 32-bit word used for `choice`)."""
    b = philox4x32_10((var, cycle, 0, 0), (seed & MASK, (seed >> 32) & MASK))
    u = ((b[0] >> 5) * 67108864 + (b[1] >> 6)) / 9007199254740992.0
    return u, b[2]


def choice_index(word: int, n: int) -> int:
    return (word * n) >> 32
