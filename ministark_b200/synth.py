"""Synthetic workloads shared by bench.py and the tests: deterministic traces and the
SynthLinear(W, T=W) AIR  f_{W+j} = w_N * f_j - f_{(j+1) mod W}  (same form as tests/e2e_goldilocks.rs:48-51)."""
import numpy as np

P = {0: 2**64 - 2**32 + 1, 1: 2013265921}
MASK = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(MASK)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(MASK)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(MASK)
    return z ^ (z >> np.uint64(31))


def synth_trace(field: int, n: int, w: int, seed: int = 0x5EED000000000000) -> np.ndarray:
    """row-major [n, w] canonical elements: splitmix64(seed ^ (row*w + col)) mod p"""
    with np.errstate(over="ignore"):
        idx = np.arange(n * w, dtype=np.uint64) ^ np.uint64(seed & MASK)
        v = splitmix64(idx) % np.uint64(P[field])
    return v.reshape(n, w).astype(np.uint64 if field == 0 else np.uint32)


def root_of_unity(field: int, log_n: int) -> int:
    p = P[field]
    gen, s = (7, 32) if field == 0 else (440564289, 27)
    return pow(pow(gen, (p - 1) >> s, p), 1 << (s - log_n), p)


def synth_linear_matrix(field: int, n: int, w: int) -> np.ndarray:
    p = P[field]
    om = root_of_unity(field, n.bit_length() - 1)
    m = np.zeros((w, w), dtype=object)
    for j in range(w):
        m[j, j] = (m[j, j] + om) % p
        m[j, (j + 1) % w] = (m[j, (j + 1) % w] - 1) % p
    return m.astype(np.uint64 if field == 0 else np.uint32)


class SynthAir:
    """pyref-compatible AIR over a given row-major trace and linear constraint matrix."""

    def __init__(self, pyref, field: int, trace_rm: np.ndarray, matrix: np.ndarray, steps: int, constants=None):
        self.R, self.F = pyref, pyref.FIELDS[field]
        self.trace_rm, self.matrix, self.steps = trace_rm, matrix, steps
        self.constants = [0] * len(matrix) if constants is None else [int(c) for c in constants]  # affine closures: + c_t

    def trace(self, _witness=None):
        R, F = self.R, self.F
        n, w = self.trace_rm.shape
        t = R.TraceTable(F, self.steps, w, padding=0)
        assert t.length == n
        t.data = [int(v) for v in self.trace_rm.reshape(-1)]
        for row, cst in zip(self.matrix, self.constants):
            coef = [int(c) for c in row]

            def f(P, coef=coef, cst=cst):
                acc = [cst % F.p] if cst % F.p else []
                for c, poly in zip(coef, P):
                    if c:
                        acc = R.poly_add(F, acc, R.poly_scale(F, poly, c))
                return acc

            t.add_transition_constrain(f, coef)
        return t
