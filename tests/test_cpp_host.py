"""The C++ host mirror (include/ministark.hpp) and its e2e example (examples/e2e_fibonacci.cpp = tests/e2e_goldilocks.rs +
tests/e2e_babybear.rs of the reference): builds with g++ against libministark.so, its host-side half (trace, test_rng padding,
affine form of the closures, derived parameters) equals the Python mirror's and the committed goldens, and without a GPU the
prover half fails loudly (there is no CPU fallback).  examples/multi_rank_local.cpp (the multi-GPU prover driven by host threads
through ms_comm_init_local / ms_stark_prove_multi, no Python in the process) builds and fails the same way.  The GPU half is
tests/test_gpu_cpp_host.py."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "e2e_proofs.json")


def build_example(tmp_path, name: str = "e2e_fibonacci") -> str:
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    from ministark_b200 import build

    lib = build.build()
    exe = str(tmp_path / name)
    cmd = [gxx, "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", name + ".cpp"),
           "-L" + os.path.dirname(lib), "-lministark", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_cpp_mirror_host_side_equals_the_python_mirror(tmp_path):
    from ministark_b200.air import DensePolynomial, TraceTable, padding_value
    from ministark_b200.field import FIELDS

    exe = build_example(tmp_path)
    r = subprocess.run([exe, "host"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    golden = json.load(open(GOLDEN))
    lines = {l.split()[0]: l.split() for l in r.stdout.splitlines() if l.strip()}
    for fid, steps in ((0, 9), (1, 7)):
        F = FIELDS[fid]
        tok = lines[F.name]
        val = lambda key, n=1: [int(x) for x in tok[tok.index(key) + 1: tok.index(key) + 1 + n]]
        # the Python mirror of the same claim
        t = TraceTable(F, steps, 3)
        a, b = 1, 2
        c = (a + b) % F.p
        for i in range(steps):
            t.add_row(i, [a, b, c])
            a, b = b, c
            c = (a + b) % F.p
        om = DensePolynomial(F, [t.omega])
        t.add_transition_constrain(lambda P: P[0].clone() * om - P[1].clone())
        t.add_transition_constrain(lambda P: P[0].clone() * om - P[1].clone())
        t.add_transition_constrain(lambda P: P[2].clone() - P[0].clone() - P[1].clone())
        m, consts = t.affine_form()
        g = golden[F.name]
        assert val("padding_value") == [padding_value(F)] == [int(g["padding_value"])]
        assert val("length") == [t.length] and val("width") == [3] and val("omega") == [t.omega]
        assert val("constrain_number") == [t.constrain_number()]
        assert val("rounds") == [g["rounds"]] and val("constrain_queries") == [g["constrain_queries"]] and val("fri_queries") == [g["fri_queries"]]
        assert val("matrix", 9) == [int(x) for x in np.asarray(m).reshape(-1)]
        assert val("constants", 3) == [int(x) for x in consts]
        assert val("last_row", 3) == [int(x) for x in t.data[steps - 1]]


def test_cpp_prover_fails_loudly_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the prover half runs in tests/test_gpu_cpp_host.py")
    exe = build_example(tmp_path)
    r = subprocess.run([exe, "prove", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr  # ministark::Error, not a silent CPU path
    assert not os.path.exists(tmp_path / "Goldilocks.proof")


def test_cpp_multi_rank_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch

    exe = build_example(tmp_path, "multi_rank_local")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the example runs in tests/test_gpu_cpp_host.py")
    r = subprocess.run([exe, "2", "10", "4"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr
