"""The reference's e2e tests through the C++ host mirror on the GPU (examples/e2e_fibonacci.cpp over include/ministark.hpp
and the C ABI, no Python in the prover's process): the proofs equal the committed goldens of the oracle byte for byte
(tests/golden/e2e_proofs.json), Stark::verify accepts them (also strict) and rejects a corrupted opening."""
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
