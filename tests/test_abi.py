"""CPU-only checks of the drop-in boundary: libministark.so loads without a GPU, exports every symbol
include/ministark.h declares (and nothing is declared that the ctypes binding does not know), the
host-side parameter derivation (StarkConfig::new, src/starks.rs:268-332) answers without a device, the
host mirror of the reference API behaves like src/air.rs, and the product never reaches into oracle/."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "ministark.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ministark_b200 import _lib
    from ministark_b200 import build as b

    b.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ministark.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert lib.ms_version() >= 1


def test_no_compute_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from ministark_b200 import Context, MiniStarkError

    with pytest.raises(MiniStarkError):
        Context(0)  # no CPU fallback


@pytest.mark.parametrize("field,args,want", [
    (0, (20, 4, 129), (1, 3)), (0, (128, 4, 129), (3, 19)), (0, (256, 4, 513), (5, 32)),  # src/starks.rs:348-374
])
def test_stark_derive_matches_reference_unit_tests(field, args, want, pyref):
    from ministark_b200 import _lib

    lib = _lib.load()
    sec, blow, steps = args
    p = _lib.StarkParams(sec, blow, steps, 6, 2)
    r, q, fq = C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert lib.ms_stark_derive(field, C.byref(p), C.byref(r), C.byref(q), C.byref(fq)) == 0
    assert (q.value, fq.value) == want
    cfg = pyref.StarkConfig(pyref.FIELDS[field], sec, blow, steps, 6)
    assert (r.value, q.value, fq.value) == (cfg.rounds, cfg.constrain_queries, cfg.fri_config.queries)


def test_stark_derive_rejects_low_security():
    from ministark_b200 import _lib

    lib = _lib.load()
    p = _lib.StarkParams(19, 4, 9, 6, 2)  # starks.rs:317-320 panics below 20 bits
    assert lib.ms_stark_derive(0, C.byref(p), None, None, None) != 0


def test_merkle_node_count_matches_reference_unit_tests():
    from ministark_b200 import _lib

    lib = _lib.load()
    # merkle.rs:399-419: 16 leaves with (lpn, k) = (2,2),(4,2),(4,4),(16,16) -> 31/23/21/17 nodes incl. leaves;
    # the digest count is that minus the 16 raw leaves
    for lpn, k, total in ((2, 2, 31), (4, 2, 23), (4, 4, 21), (16, 16, 17)):
        assert lib.ms_merkle_node_count(16 // lpn, k) == total - 16
    assert lib.ms_merkle_node_count(8, 4) == 0  # not full (merkle.rs:384-396 panics)


def test_trace_table_mirror_matches_air_rs(pyref):
    """air.rs:245-300: padded sizes 3->4, 4->8, 5->8 rows; padding cells equal and non-zero; linear matrix
    of the e2e AIR's closures."""
    from ministark_b200.air import DensePolynomial, TraceTable
    from ministark_b200.field import FIELDS

    F = FIELDS[0]
    for steps, rows in ((3, 4), (4, 8), (5, 8), (9, 16)):
        t = TraceTable(F, steps, 3)
        assert t.data.shape == (rows, 3)
        pad = t.data[steps:]
        assert (pad == pad[0, 0]).all() and int(pad[0, 0]) != 0
        rt = pyref.TraceTable(pyref.FIELDS[0], steps, 3)
        assert [int(v) for v in t.data.reshape(-1)] == rt.data
    t = TraceTable(F, 9, 3)
    om = DensePolynomial(F, [t.omega])
    t.add_transition_constrain(lambda P: P[0].clone() * om - P[1].clone())
    t.add_transition_constrain(lambda P: P[2].clone() - P[0].clone() - P[1].clone())
    m = t.linear_matrix()
    p = F.p
    assert m.tolist() == [[t.omega, p - 1, 0], [p - 1, p - 1, 1]]
    assert t.constrain_number() == 5
    # closures with an additive constant polynomial (an affine AIR: provable by the reference, an affine combination still
    # has at most N coefficients): the host mirror separates the matrix from the constants
    t.add_transition_constrain(lambda P: P[1].clone() * DensePolynomial(F, [3]) + DensePolynomial(F, [p - 4]))
    with pytest.raises(ValueError):
        t.linear_matrix()
    m2, c2 = t.affine_form()
    assert m2.tolist() == [[t.omega, p - 1, 0], [p - 1, p - 1, 1], [0, 3, 0]] and c2.tolist() == [0, 0, p - 4]
    t.add_transition_constrain(lambda P: P[0].clone() * P[1].clone())  # a product: degree >= N, starks.rs:119 panics
    with pytest.raises(ValueError):
        t.affine_form()


@pytest.mark.parametrize("name,steps", [("Goldilocks", 9), ("BabyBear", 7)])
def test_proof_dump_parser_round_trips_the_oracle_proofs(name, steps, pyref):
    """StarkProof.from_bytes (the host mirror's parser of the canonical dump) against the
    restated reference prover on the two reference e2e configurations (tests/e2e_goldilocks.rs,
    tests/e2e_babybear.rs): same bytes as the committed golden hashes, every field recovered."""
    import hashlib
    import json

    from ministark_b200.starks import StarkProof

    F = getattr(pyref, name)
    claim = pyref.FibonacciClaim(F, steps)
    trace = claim.trace(2)
    cfg = pyref.StarkConfig(F, 20, 2, trace.step_number(), trace.constrain_number())
    ref = pyref.Stark(cfg).prove(claim, 2)
    raw = pyref.serialize_proof(F, ref)
    with open(os.path.join(ROOT, "tests", "golden", "e2e_proofs.json")) as fh:
        golden = json.load(fh)[name]
    assert len(raw) == golden["proof_len"] and hashlib.sha256(raw).hexdigest() == golden["proof_sha256"]
    got = StarkProof.from_bytes(raw)
    assert got.arthur == ref.arthur and got.trace_commit == ref.trace_commit
    assert got.trace_commit.hex() == golden["trace_commit"]
    assert got.constrain_trace_commit == ref.constrain_trace_commit
    norm = lambda e: tuple(int(c) for c in (e if isinstance(e, (tuple, list)) else (e,)))
    assert [[norm(e) for e in row] for row in got.constrain_queries] == [[norm(e) for e in row] for row in ref.constrain_queries]
    assert [norm(e) for e in got.validity_queries] == [norm(e) for e in ref.validity_queries]
    assert len(got.fri_proof.points) == len(ref.fri_proof.points) == golden["rounds"] - 1
    assert all(len(r) == golden["fri_queries"] for r in got.fri_proof.quotients)
    for gr, rr in zip(got.fri_proof.quotients, ref.fri_proof.quotients):
        for gq, rq in zip(gr, rr):
            assert [norm(e) for e in gq] == [norm(e) for e in rq]
    # tampered counts end in an AssertionError (or parse, where the word was data), never in an unbounded loop or allocation:
    # every u64 of the first kilobyte after the commitments blown up, and the (Q, C) pair of the openings as (2^60, 0)
    import struct

    at = 24 + len(ref.arthur) + 64
    for off in range(at, min(at + 1024, len(raw) - 8), 8):
        bad = bytearray(raw)
        bad[off:off + 8] = struct.pack("<Q", (1 << 60) + 1)
        try:
            StarkProof.from_bytes(bytes(bad))  # the word was data, not a count
        except AssertionError:
            pass
    bad = bytearray(raw)
    bad[at:at + 16] = struct.pack("<QQ", 1 << 60, 0)
    with pytest.raises(AssertionError):
        StarkProof.from_bytes(bytes(bad))


def test_rust_shim_binds_every_declared_symbol():
    """shim/src/gpu/ffi.rs (the Rust `extern "C"` block a maintainer of the reference adds; generated from the header by
    tools/gen_rust_ffi.py) names every entry point include/ministark.h declares, with the same number of arguments, and
    is up to date with the header."""
    import subprocess
    import sys

    with open(os.path.join(ROOT, "shim", "src", "gpu", "ffi.rs")) as fh:
        rs = fh.read()
    with open(os.path.join(ROOT, "include", "ministark.h")) as fh:
        hdr = re.sub(r"/\*.*?\*/", " ", fh.read(), flags=re.S)
    for name in _declared_symbols():
        m = re.search(rf"pub fn {name}\(([^)]*)\)", rs)
        assert m, f"{name} is not bound in shim/src/gpu/ffi.rs"
        c = re.search(rf"\b{name}\s*\(([^;{{}}]*?)\)\s*;", hdr)
        c_args = [a for a in c.group(1).split(",") if a.strip() and a.strip() != "void"]
        r_args = [a for a in m.group(1).split(",") if a.strip()]
        assert len(c_args) == len(r_args), (name, c_args, r_args)
    # regenerating changes nothing
    before = rs
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py")], stdout=subprocess.DEVNULL)
    with open(os.path.join(ROOT, "shim", "src", "gpu", "ffi.rs")) as fh:
        assert fh.read() == before, "shim/src/gpu/ffi.rs is stale: run python tools/gen_rust_ffi.py"


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "ministark_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                code = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("#", "//", "*", '"""')))
                assert not re.search(r"^\s*(from|import)\s+oracle\b", code, flags=re.M), f
                assert "liboracle" not in code, f


def test_context_free_entry_points_survive_edge_values():
    """ms_merkle_node_count, ms_stark_derive, ms_stark_proof_bound, ms_shard_plan take no context and run anywhere: zeros, ones,
    non powers of two, 2^32, 2^63, 2^64 - 1, bad field ids and negative ranks come back as a value or an error code (no
    division by zero, no endless loop)."""
    import ctypes as C
    import faulthandler

    from ministark_b200 import _lib
    from ministark_b200._lib import StarkParams

    faulthandler.enable()
    lib = _lib.load()
    vals = [0, 1, 2, 3, 4, 7, 8, 16, 20, 63, 64, 100, 1 << 20, (1 << 32) - 1, 1 << 32, 1 << 63, (1 << 64) - 1]
    for n in vals:
        for k in vals:
            lib.ms_merkle_node_count(n, k)
    # digests for the leaf-group counts of merkle.rs:399-419 (16 leaves under TWO / TWO_FOUR / FOUR / SIXTEEN); 12 groups 4-ary: not full
    assert [lib.ms_merkle_node_count(g, k) for g, k in ((8, 2), (4, 2), (4, 4), (1, 16), (12, 4))] == [15, 7, 5, 1, 0]
    r, cq, fq = C.c_uint64(), C.c_uint64(), C.c_uint64()
    for f in (0, 1, 2, -1):
        for sec in vals:
            for b in vals:
                for steps in vals:
                    p = StarkParams(sec, b, steps, 8, 2)
                    rc = lib.ms_stark_derive(f, C.byref(p), C.byref(r), C.byref(cq), C.byref(fq))
                    assert rc != 0 or (f in (0, 1) and sec >= 20)
                    for n in (0, 1, 16, 1 << 22, 1 << 63, (1 << 64) - 1):
                        lib.ms_stark_proof_bound(f, C.byref(p), n, 32)
    a, b_, per, left = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    for cols in vals:
        for groups in vals:
            for k in (0, 1, 2, 3, 4, 1 << 63):
                for world in (-1, 0, 1, 2, 3, 8, 64, 1 << 20):
                    for rank in (-1, 0, 1, 7):
                        rc = lib.ms_shard_plan(cols, groups, k, world, rank, C.byref(a), C.byref(b_), C.byref(per), C.byref(left))
                        if rc == 0:
                            assert 0 <= rank < world and a.value <= b_.value <= cols


def test_chacha_known_answers_pin_the_test_rng_core():
    """ark_std::test_rng() is rand 0.8's StdRng = ChaCha12.  The block function of the host mirror
    (ministark_b200/air.py, behind `padding_value`) against published known answers: all-zero key / IV / counter, first
    keystream block at 8, 12 and 20 rounds (draft-strombergson-chacha-test-vectors TC1; the 20-round one is also the
    all-zero vector of RFC 7539 section 2.3.2's construction) -- and the oracle's copy gives the same blocks."""
    import struct

    from ministark_b200.air import _chacha_block
    from oracle import pyref

    want = {
        8: "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42",
        12: "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be",
        20: "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586",
    }
    for rounds, hexstr in want.items():
        assert struct.pack("<16I", *_chacha_block([0] * 8, 0, rounds)).hex() == hexstr
        assert struct.pack("<16I", *pyref._chacha_block([0] * 8, 0, rounds)).hex() == hexstr
    # the 64-bit block counter occupies words 12-13 (rand_chacha): block 1 differs from block 0 and from a nonce change
    assert _chacha_block([0] * 8, 1, 12) != _chacha_block([0] * 8, 0, 12)
