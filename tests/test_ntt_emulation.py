"""CPU check of the CUDA NTT's index maps: csrc/ntt.cuh's per-thread round bodies (generic k_ntt_tile and
the compile-time-shaped k_ntt_fixed) are __host__ __device__, so tests/emul/emul_ntt.cu runs them thread by
thread on the host -- tile geometry, round split, shared-memory swizzle, bit-reversed gathers, two-pass
index maps (in place and through a temporary), coset-folded twiddles and the inter-pass factor table --
against a naive Horner evaluation, for both fields; the Goldilocks register block with power-of-two twiddles
(gl_shift_dft, both directions) and x * 2^S for every S in 1..95 are also checked on their own.  (The PTX arithmetic itself is covered on the GPU by
test_butterfly_arithmetic_selftest.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ntt_tile_emulation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "emul_ntt")
    src = os.path.join(ROOT, "tests", "emul", "emul_ntt.cu")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]
    assert "MISMATCH" not in r.stdout
    used = [l for l in r.stdout.splitlines() if l.startswith("fixed-shape tiles used")]
    assert used and int(used[0].split(":")[1]) > 20  # the specialised kernel bodies were exercised
