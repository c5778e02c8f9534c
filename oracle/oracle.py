"""ctypes front-end of the C oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (ministark_b200/) never does.  Every function is a thin wrapper of the
C restatement in oracle.c / oracle_impl.inc, which cite the reference lines they follow.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")

GL, BB = 0, 1
MODULUS = {GL: 2**64 - 2**32 + 1, BB: 2013265921}
EXT_DEGREE = {GL: 2, BB: 4}


def build(force: bool = False) -> str:
    """Compile liboracle.so with the Makefile next to this file (gcc only)."""
    srcs = [os.path.join(_DIR, f) for f in ("oracle.c", "oracle_impl.inc", "prover.inc", "transcript.inc", "fields.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s)
    ):
        subprocess.check_call(["make", "-C", _DIR, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u64, i32, vp = C.c_uint64, C.c_int, C.c_void_p
        sig = {
            "or_sha256": (None, [vp, u64, vp]),
            "or_set_zero_display": (None, [i32]),
            "or_leaf_string": (u64, [vp, i32, vp]),
            "or_merkle": (C.c_int64, [vp, i32, u64, u64, u64, vp, vp]),
            "or_keccak_f1600": (None, [vp]),
            "or_modulus": (u64, [i32]),
            "or_ext_degree": (i32, [i32]),
            "or_root_of_unity": (u64, [i32, C.c_uint]),
            "or_fmul": (u64, [i32, u64, u64]),
            "or_fpow": (u64, [i32, u64, u64]),
            "or_finv": (u64, [i32, u64]),
            "or_ext_mul": (None, [i32, vp, vp, vp]),
            "or_ntt": (None, [i32, vp, u64, i32]),
            "or_eval_domain_naive": (None, [i32, vp, u64, u64, u64, vp]),
            "or_trace_polys": (None, [i32, vp, u64, u64, vp]),
            "or_coset_lde": (None, [i32, vp, u64, u64, u64, u64, vp, i32]),
            "or_mix": (None, [i32, vp, u64, u64, u64, vp]),
            "or_eval_base_at_ext": (None, [i32, vp, u64, vp, vp]),
            "or_eval_ext_at_ext": (None, [i32, vp, u64, vp, vp]),
            "or_fri_codeword": (None, [i32, vp, u64, u64, vp]),
            "or_fri_fold": (u64, [i32, vp, u64, vp, vp, vp, vp]),
            "or_fri_query_quotient": (u64, [i32, vp, u64, u64, u64, vp, vp, vp]),
            "or_merkle_threads": (C.c_int64, [vp, i32, u64, u64, u64, vp, vp, i32]),
            "or_set_bridge_masks": (None, [i32, i32, i32]),
            "or_set_leftover_mode": (None, [i32]),
            "or_stark_derive": (i32, [i32, vp, vp, vp, vp]),
            "or_transcript_squeeze_test": (C.c_int64, [i32, u64, u64, u64, vp, u64, vp]),
            "or_stark_prove": (C.c_int64, [i32, vp, vp, u64, u64, vp, vp, u64, vp, u64, i32, vp]),
            "or_stark_verify": (i32, [i32, vp, vp, u64, u64, vp, u64, i32, vp]),
            "or_derive_constrains": (None, [i32, vp, u64, u64, vp, vp, u64, vp, i32]),
            "or_stark_proof_bound": (u64, [i32, vp, u64, u64]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def sha256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().or_sha256(data, len(data), out)
    return out.raw


def set_zero_display(empty: bool) -> None:
    lib().or_set_zero_display(int(empty))


def leaf_string(elem, deg: int = 1) -> bytes:
    e = _u64(elem).reshape(-1)
    out = C.create_string_buffer(512)
    n = lib().or_leaf_string(_p(e), deg, out)
    return out.raw[:n]


def merkle(data, lpn: int, k: int = 2, deg: int = 1, want_nodes: bool = False):
    """MerkleTree::new (merkle.rs:81-148). data: flat element array (n_elems*deg uint64).
    Returns root bytes, or (root, nodes[n_nodes,32]) with want_nodes."""
    d = _u64(data).reshape(-1)
    n_elems = d.size // deg
    root = C.create_string_buffer(32)
    nodes = None
    if want_nodes:
        n1 = n_elems // lpn
        total, lv = 0, n1
        while lv >= 1:
            total += lv
            if lv == 1:
                break
            lv //= k
        nodes = np.zeros((total, 32), dtype=np.uint8)
    r = lib().or_merkle(_p(d), deg, n_elems, lpn, k, _p(nodes) if nodes is not None else None, root)
    if r < 0:
        raise ValueError("tree is not full / bad shape (merkle.rs:93-104 panics)")
    return (root.raw, nodes) if want_nodes else root.raw


def root_of_unity(field: int, log_n: int) -> int:
    return int(lib().or_root_of_unity(field, log_n))


def fpow(field, a, e):
    return int(lib().or_fpow(field, a, e))


def finv(field, a):
    return int(lib().or_finv(field, a))


def ext_mul(field, a, b):
    a, b = _u64(a), _u64(b)
    r = np.zeros(EXT_DEGREE[field], dtype=np.uint64)
    lib().or_ext_mul(field, _p(a), _p(b), _p(r))
    return r


def ntt(field: int, a, inverse: bool = False) -> np.ndarray:
    a = _u64(a).copy()
    lib().or_ntt(field, _p(a), a.size, int(inverse))
    return a


def eval_domain_naive(field, coef, n, offset=1) -> np.ndarray:
    coef = _u64(coef)
    out = np.zeros(n, dtype=np.uint64)
    lib().or_eval_domain_naive(field, _p(coef), coef.size, n, offset, _p(out))
    return out


def trace_polys(field, trace_rm) -> np.ndarray:
    """air.rs:147-160. trace_rm: [N, W] row-major -> [W, N] coefficient vectors."""
    t = _u64(trace_rm)
    N, W = t.shape
    out = np.zeros((W, N), dtype=np.uint64)
    lib().or_trace_polys(field, _p(t), N, W, _p(out))
    return out


def coset_lde(field, polys_cm, L: int, shift: int, threads: int = 1) -> np.ndarray:
    """starks.rs:82-91. polys_cm: [C, N] coefficients -> row-major [L, C] evaluations on shift*<w_L>."""
    pc = _u64(polys_cm)
    Cn, N = pc.shape
    out = np.zeros((L, Cn), dtype=np.uint64)
    lib().or_coset_lde(field, _p(pc), N, Cn, L, shift, _p(out), threads)
    return out


def mix(field, polys_cm, r: int) -> np.ndarray:
    pc = _u64(polys_cm)
    Cn, N = pc.shape
    out = np.zeros(N, dtype=np.uint64)
    lib().or_mix(field, _p(pc), N, Cn, r, _p(out))
    return out


def eval_base_at_ext(field, coef, z) -> np.ndarray:
    coef, z = _u64(coef), _u64(z)
    out = np.zeros(EXT_DEGREE[field], dtype=np.uint64)
    lib().or_eval_base_at_ext(field, _p(coef), coef.size, _p(z), _p(out))
    return out


def eval_ext_at_ext(field, coef, z) -> np.ndarray:
    D = EXT_DEGREE[field]
    coef, z = _u64(coef).reshape(-1, D), _u64(z)
    out = np.zeros(D, dtype=np.uint64)
    lib().or_eval_ext_at_ext(field, _p(coef), coef.shape[0], _p(z), _p(out))
    return out


def fri_codeword(field, poly, n: int) -> np.ndarray:
    D = EXT_DEGREE[field]
    poly = _u64(poly).reshape(-1, D)
    out = np.zeros((n, D), dtype=np.uint64)
    lib().or_fri_codeword(field, _p(poly), poly.shape[0], n, _p(out))
    return out


def fri_fold(field, poly, z, alpha):
    """fri.rs:89-101. Returns (d[2, D], next_poly[n', D])."""
    D = EXT_DEGREE[field]
    poly, z, alpha = _u64(poly).reshape(-1, D), _u64(z), _u64(alpha)
    d = np.zeros((2, D), dtype=np.uint64)
    nxt = np.zeros((max(poly.shape[0], 1), D), dtype=np.uint64)
    n = lib().or_fri_fold(field, _p(poly), poly.shape[0], _p(z), _p(alpha), _p(d), _p(nxt))
    return d, nxt[:n].copy()


def fri_query_quotient(field, poly, x1: int, x2: int, y1, y2) -> np.ndarray:
    D = EXT_DEGREE[field]
    poly, y1, y2 = _u64(poly).reshape(-1, D), _u64(y1), _u64(y2)
    q = np.zeros((max(poly.shape[0], 1), D), dtype=np.uint64)
    n = lib().or_fri_query_quotient(field, _p(poly), poly.shape[0], x1, x2, _p(y1), _p(y2), _p(q))
    return q[:n].copy()


# ---------------------------------------------------------------------------------------- whole prover
class StarkParams(C.Structure):
    """StarkConfig::new arguments (starks.rs:268-273) + the inner_children extension; same layout as ms_stark_params."""

    _fields_ = [(n, C.c_uint64) for n in ("security_bits", "blowup_factor", "steps", "trace_columns", "inner_children")]


STAGES = ("trace_commit", "intt+constraints", "lde", "lde_commit", "mix+open", "fri_commit_phase", "fri_query_phase", "total")
PROVE_ERRORS = {-1: "bad shape (reference panics)", -3: "transcript pattern violated", -4: "leaf is not included in the tree",
                -5: "random_shift is zero"}


def set_leftover_mode(as_published: bool) -> None:
    lib().or_set_leftover_mode(int(as_published))


def set_bridge_masks(absorb: int, squeeze: int, squeeze_end: int) -> None:
    lib().or_set_bridge_masks(absorb, squeeze, squeeze_end)


def stark_derive(field: int, security_bits: int, blowup: int, steps: int):
    p = StarkParams(security_bits, blowup, steps, 1, 2)
    r, cq, fq = C.c_uint64(), C.c_uint64(), C.c_uint64()
    if lib().or_stark_derive(field, C.byref(p), C.byref(r), C.byref(cq), C.byref(fq)):
        raise ValueError("bad STARK parameters")
    return r.value, cq.value, fq.value


def stark_prove(field: int, security_bits: int, blowup: int, steps: int, trace_columns: int, trace_rm, matrix,
                inner_children: int = 2, threads: int = 1, want_timings: bool = False, constants=None):
    """Stark::prove (starks.rs:59-169) in C: returns the canonical proof bytes (and the per-stage wall ms).
    constants: optional T additive constants (affine transition closures)."""
    t = _u64(trace_rm)
    N, W = t.shape
    m = _u64(matrix).reshape(-1, W) if np.size(matrix) else np.zeros((0, W), dtype=np.uint64)
    cst = None if constants is None else _u64(constants).reshape(m.shape[0])
    p = StarkParams(security_bits, blowup, steps, trace_columns, inner_children)
    ms = (C.c_double * len(STAGES))()
    bound = lib().or_stark_proof_bound(field, C.byref(p), N, W + m.shape[0])
    if bound == 0:
        raise ValueError(PROVE_ERRORS[-1])
    buf = np.empty(bound, dtype=np.uint8)
    got = lib().or_stark_prove(field, C.byref(p), _p(t), N, W, _p(m), _p(cst) if cst is not None else None, m.shape[0], _p(buf), bound, threads, ms)
    if got < 0:
        raise ValueError(PROVE_ERRORS.get(got, f"or_stark_prove: {got}"))
    assert got <= bound
    buf = buf[:got]
    return (buf, dict(zip(STAGES, ms))) if want_timings else buf


def stark_prove_into(field: int, params: "StarkParams", trace_rm, matrix, out: np.ndarray, threads: int = 1):
    """single pass into a caller buffer (bench: no sizing pass); returns (length, stage ms)"""
    t = _u64(trace_rm)
    N, W = t.shape
    m = _u64(matrix).reshape(-1, W)
    ms = (C.c_double * len(STAGES))()
    got = lib().or_stark_prove(field, C.byref(params), _p(t), N, W, _p(m), None, m.shape[0], _p(out), out.size, threads, ms)
    if got < 0:
        raise ValueError(PROVE_ERRORS.get(got, f"or_stark_prove: {got}"))
    return got, dict(zip(STAGES, ms))


def derive_constrains(field: int, trace_rm, matrix, threads: int = 1, constants=None) -> np.ndarray:
    """TraceTable::derive_constrains (air.rs:127-144) for a linear (affine with `constants`) AIR: [W + T, N] coefficient vectors."""
    t = _u64(trace_rm)
    N, W = t.shape
    m = _u64(matrix).reshape(-1, W)
    cst = None if constants is None else _u64(constants).reshape(m.shape[0])
    out = np.zeros((W + m.shape[0], N), dtype=np.uint64)
    lib().or_derive_constrains(field, _p(t), N, W, _p(m), _p(cst) if cst is not None else None, m.shape[0], _p(out), threads)
    return out


def stark_verify(field: int, security_bits: int, blowup: int, steps: int, trace_columns: int, constrains_cm, proof,
                 inner_children: int = 2, strict: bool = True):
    """Stark::verify (starks.rs:171-235) in C.  Returns (accepted, line of the failed check or 0)."""
    c = _u64(constrains_cm)
    Cn, N = c.shape
    pr = np.frombuffer(bytes(proof), dtype=np.uint8) if not isinstance(proof, np.ndarray) else np.ascontiguousarray(proof, dtype=np.uint8)
    p = StarkParams(security_bits, blowup, steps, trace_columns, inner_children)
    why = C.c_int(0)
    r = lib().or_stark_verify(field, C.byref(p), _p(c), N, Cn, _p(pr), pr.size, int(strict), C.byref(why))
    if r < 0:
        raise ValueError(f"malformed proof dump ({r})")
    return bool(r), why.value
