"""The reference's e2e tests through the C++ host mirror on the GPU (examples/e2e_fibonacci.cpp over include/ministark.hpp
and the C ABI, no Python in the prover's process): the proofs equal the committed goldens of the oracle byte for byte
(tests/golden/e2e_proofs.json), Stark::verify accepts them (also strict) and rejects a corrupted opening.  And the multi-GPU prover
from compiled host code (examples/multi_rank_local.cpp: ms_comm_init_local + one host thread per rank + ms_stark_prove_multi;
the ranks share the test box's one GPU): the sharded proof equals the single-context proof."""
import hashlib
import json
import os
import subprocess

import pytest

from tests.test_cpp_host import GOLDEN, build_example

pytestmark = pytest.mark.gpu


def test_cpp_e2e_fibonacci_proofs_equal_the_goldens(tmp_path):
    exe = build_example(tmp_path)
    r = subprocess.run([exe, "prove", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    golden = json.load(open(GOLDEN))
    for name in ("Goldilocks", "BabyBear"):
        line = [l for l in r.stdout.splitlines() if l.startswith(name)][0].split()
        flags = dict(zip(line[1::2], line[2::2]))
        assert flags["verify"] == "1" and flags["strict"] == "1" and flags["corrupted_rejected"] == "1"
        raw = open(os.path.join(tmp_path, name + ".proof"), "rb").read()
        g = golden[name]
        assert len(raw) == g["proof_len"] == int(flags["proof_len"])
        assert hashlib.sha256(raw).hexdigest() == g["proof_sha256"]
        alen = int.from_bytes(raw[16:24], "little")
        assert raw[24:24 + alen].hex() == g["arthur"] and int(flags["arthur_len"]) == alen
        assert raw[24 + alen:24 + alen + 32].hex() == g["trace_commit"]
        assert raw[24 + alen + 32:24 + alen + 64].hex() == g["constrain_trace_commit"]


@pytest.mark.parametrize("world,log_rows,width", [(2, 14, 8), (4, 12, 4), (3, 10, 3)])
def test_cpp_multi_rank_proof_equals_the_single_context_proof(tmp_path, world, log_rows, width):
    exe = build_example(tmp_path, "multi_rank_local")
    r = subprocess.run([exe, str(world), str(log_rows), str(width)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    flags = dict(zip(r.stdout.split()[0::2], r.stdout.split()[1::2]))
    assert flags["world"] == str(world) and flags["backend"] == "local" and flags["last_rank"] == str(world - 1)
    assert flags["identical"] == "1" and flags["proof_len"] == flags["single_len"] and int(flags["proof_len"]) > 0
