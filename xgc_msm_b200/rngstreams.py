"""L'Ecuyer's RngStreams (MRG32k3a) — the generator behind `rlecuyer 0.3-5` (renv.lock:734), which the reference's HPC
wrappers use to give every simulation its own stream — so that an injected-draw table can be filled with the SAME
uniforms an R session would draw (`xgc_inject_table()` in rshim/modules.R fills it with `runif()`), and a run driven
from R can be reproduced from this harness draw for draw ("rlecuyer-compatible stream import").

Algorithm (L'Ecuyer 1999; L'Ecuyer, Simard, Chen, Kelton 2002, published; restated here, no code taken from anywhere):
two order-3 linear recurrences modulo m1 = 2^32 - 209 and m2 = 2^32 - 22853,

    x1[n] = (1403580 x1[n-2] -  810728 x1[n-3]) mod m1
    x2[n] = ( 527612 x2[n-1] - 1370589 x2[n-3]) mod m2
    u[n]  = ((x1[n] - x2[n]) mod m1) / (m1 + 1)            (0 is mapped to m1 / (m1 + 1))

Streams are 2^127 steps apart and substreams 2^76 steps apart; the jump matrices are computed here by repeated
squaring of the one-step matrices (no tabulated constants to mistype).  Known answers (seed 12345 x 6): the second and
third uniforms are 0.3185275653 and 0.3091860155 — pinned in tests/test_oracle_cpu.py.

This is host-side tooling around the injected-draw entry point; the device's native generator stays Philox4x32-10.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

M1, M2 = 4294967087, 4294944443
A12, A13N, A21, A23N = 1403580, 810728, 527612, 1370589
NORM = 2.328306549295727688e-10            # 1 / (m1 + 1)
DEFAULT_SEED = (12345,) * 6

_A1 = ((0, 1, 0), (0, 0, 1), (M1 - A13N, A12, 0))
_A2 = ((0, 1, 0), (0, 0, 1), (M2 - A23N, 0, A21))


def _matmul(a, b, m):
    return tuple(tuple(sum(a[i][k] * b[k][j] for k in range(3)) % m for j in range(3)) for i in range(3))


def _matpow2(a, e, m):
    """a^(2^e) mod m"""
    for _ in range(e):
        a = _matmul(a, a, m)
    return a


def _matvec(a, v, m):
    return [sum(a[i][k] * v[k] for k in range(3)) % m for i in range(3)]


_JUMP = {}


def _jump(e):
    if e not in _JUMP:
        _JUMP[e] = (_matpow2(_A1, e, M1), _matpow2(_A2, e, M2))
    return _JUMP[e]


def check_seed(seed: Sequence[int]):
    s = [int(v) for v in seed]
    if len(s) != 6 or any(v < 0 for v in s) or any(v >= M1 for v in s[:3]) or any(v >= M2 for v in s[3:]) \
            or not any(s[:3]) or not any(s[3:]):
        raise ValueError("MRG32k3a seed: six integers, first three < m1 and not all 0, last three < m2 and not all 0")
    return s


class RngStream:
    """One stream.  `RngStream.create_streams(n, package_seed)` returns n successive streams exactly like n calls of
    `.lec.CreateStream` after `.lec.SetPackageSeed(package_seed)`."""

    def __init__(self, seed: Sequence[int] = DEFAULT_SEED):
        self.ig = check_seed(seed)          # start of the stream
        self.bg = list(self.ig)             # start of the current substream
        self.cg = list(self.ig)             # current state

    @staticmethod
    def create_streams(n: int, package_seed: Sequence[int] = DEFAULT_SEED) -> List["RngStream"]:
        out, s = [], check_seed(package_seed)
        a1, a2 = _jump(127)
        for _ in range(n):
            out.append(RngStream(s))
            s = _matvec(a1, s[:3], M1) + _matvec(a2, s[3:], M2)
        return out

    def reset_start_stream(self):
        self.bg = list(self.ig); self.cg = list(self.ig)

    def reset_next_substream(self):
        a1, a2 = _jump(76)
        self.bg = _matvec(a1, self.bg[:3], M1) + _matvec(a2, self.bg[3:], M2)
        self.cg = list(self.bg)

    def advance(self, n: int):
        """Jump n >= 0 steps ahead in O(log n)."""
        b1, b2, p1, p2 = _A1, _A2, None, None
        while n:
            if n & 1:
                p1 = b1 if p1 is None else _matmul(b1, p1, M1)
                p2 = b2 if p2 is None else _matmul(b2, p2, M2)
            b1, b2 = _matmul(b1, b1, M1), _matmul(b2, b2, M2)
            n >>= 1
        if p1 is not None:
            self.cg = _matvec(p1, self.cg[:3], M1) + _matvec(p2, self.cg[3:], M2)

    def rand_u01(self) -> float:
        c = self.cg
        p1 = (A12 * c[1] - A13N * c[0]) % M1
        c[0], c[1], c[2] = c[1], c[2], p1
        p2 = (A21 * c[5] - A23N * c[3]) % M2
        c[3], c[4], c[5] = c[4], c[5], p2
        return ((p1 - p2) if p1 > p2 else (p1 - p2 + M1)) * NORM

    def uniform(self, n: int) -> np.ndarray:
        """The next n uniforms (vectorised over the two recurrences with int64 arithmetic reduced step by step)."""
        out = np.empty(n)
        c = self.cg
        x10, x11, x12, x20, x21, x22 = (int(v) for v in c)
        for i in range(n):
            p1 = (A12 * x11 - A13N * x10) % M1
            x10, x11, x12 = x11, x12, p1
            p2 = (A21 * x22 - A23N * x20) % M2
            x20, x21, x22 = x21, x22, p2
            out[i] = ((p1 - p2) if p1 > p2 else (p1 - p2 + M1)) * NORM
        self.cg = [x10, x11, x12, x20, x21, x22]
        return out


def inject_table(stream: RngStream, n_replicates: int, stride: int) -> np.ndarray:
    """[n_replicates, stride] injected-draw table filled, row after row, with the stream's next uniforms — the table
    `runif(n_replicates * stride)` would give an R session drawing from the same rlecuyer stream.  Use with
    abi.make_inject(); u = 0 cannot occur (MRG32k3a never returns 0 or 1)."""
    return stream.uniform(n_replicates * stride).reshape(n_replicates, stride)
