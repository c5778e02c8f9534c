"""The PRODUCT's Fiat-Shamir transcript on a CPU-only box: csrc/transcript.hpp (the nimue IOPattern / Merlin / DigestBridge<Sha256>
that the prover inside libministark.so walks) is host code, so tests/emul/host_transcript.cu runs the
prover's whole sequence of transcript operations (src/starks.rs:59-169, src/fri.rs:64-189) with scripted absorb data and prints
every challenge; they must equal the oracle's (oracle/pyref.py), byte for byte, for both fields and both leftover modes.  On the
GPU the same code is covered through whole-proof equality; this is the one piece of the product every proof byte after the first
root depends on, and it needs no device."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("transcript") / "host_transcript")
    src = os.path.join(ROOT, "tests", "emul", "host_transcript.cu")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-I" + os.path.join(ROOT, "include"),
                        "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


@pytest.mark.parametrize("published", [True, False])
@pytest.mark.parametrize("field,rounds,cq,fq", [(0, 5, 3, 10), (1, 4, 7, 13), (0, 23, 2, 13), (1, 2, 1, 1)])
def test_product_transcript_equals_the_oracle(harness, pyref, field, rounds, cq, fq, published):
    R = pyref
    F = R.Goldilocks if field == 0 else R.BabyBear
    r = subprocess.run([harness, str(field), str(rounds), str(cq), str(fq), "1" if published else "0"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    got = {}
    for line in r.stdout.splitlines():
        k, _, v = line.partition(" ")
        got.setdefault(k, []).append(v)
    R.LEFTOVER_AS_PUBLISHED = published
    try:
        io = R.new_stark_iopattern(F, rounds, cq, fq, "\U0001F43A")
        assert got["io"] == [io.as_bytes().hex()]
        assert got["tag"] == [R.nimue_tag(io.as_bytes()).hex()]
        t = R.Transcript(F, io)
        D = F.ext_degree
        flat = lambda xs: " ".join(str(c) for x in xs for c in x)
        t.add_bytes(bytes(range(32)))
        assert got["shift"] == [str(t.challenge_base())]
        t.add_bytes(bytes(range(32, 64)))
        assert got["r"] == [str(t.challenge_base())]
        assert got["queries"] == [flat(t.challenge_ext_many(cq))]
        zs, alphas = [], []
        for i in range(rounds - 1):
            zs.append(flat([t.challenge_ext()]))
            t.add_bytes(bytes([i] * (2 * D * F.base_bytes)))
            alphas.append(flat([t.challenge_ext()]))
            t.add_bytes(bytes([0x40 + i] * 32))
        assert got.get("z", []) == zs and got.get("alpha", []) == alphas
        assert got["betas"] == [t.challenge_bytes(8 * fq).hex()]
        assert got["arthur"] == [bytes(t.transcript).hex()]
        assert got["extra_squeeze_refused"] == ["1"]
        with pytest.raises(R.IOPatternError):
            t.challenge_bytes(1)
    finally:
        R.LEFTOVER_AS_PUBLISHED = True
    # StarkConfig::new's derivation inside the same header (src/starks.rs:268-310) against the oracle's
    cfg = R.StarkConfig(F, 100, 4, 1023, 8)
    assert got["derived"] == [f"{cfg.rounds} {cfg.constrain_queries} {cfg.fri_config.queries}"]
