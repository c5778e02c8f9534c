"""Prover-level parity for DENSE constraint matrices (more than 4 non-zero entries per row): these take k_linear_constraints and
the transform of all C columns instead of the sparse kernels / the LDE by linearity that every other prover test exercises
(csrc/prover.cuh, `sparse_rows`).  Written after the round's last GPU second was spent, so it has not run on a B200 yet: it sits
in the last test file on purpose (a failure here cannot hide another test's result behind `-x`).  Same structure as
tests/test_gpu_prove.py::test_affine_constraints_match_the_oracle, which passed on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GL, BB = 0, 1


@pytest.mark.parametrize("field", [GL, BB])
@pytest.mark.parametrize("log_n,w,blowup,with_consts", [(10, 8, 4, False), (9, 6, 8, True)])
def test_dense_constraint_matrix_matches_the_oracle(field, log_n, w, blowup, with_consts, oracle):
    from ministark_b200 import Context
    from ministark_b200._lib import StarkParams
    from ministark_b200.synth import synth_trace

    n = 1 << log_n
    p = 2**64 - 2**32 + 1 if field == GL else 2013265921
    rng = np.random.default_rng(1234 + log_n)
    tr = synth_trace(field, n, w)
    mat = (rng.integers(1, 2**62, size=(w, w), dtype=np.uint64) % np.uint64(p - 1) + np.uint64(1)).astype(np.uint64)  # every entry non-zero
    assert (mat != 0).all() and w > 4
    cst = np.array([(7 * i + 1) % p for i in range(w)], dtype=np.uint64) if with_consts else None
    params = StarkParams(40, blowup, n - 1, 2 * w, 2)
    want = oracle.stark_prove(field, 40, blowup, n - 1, 2 * w, tr, mat, threads=4, constants=cst).tobytes()
    ctx = Context(field)
    try:
        raw = ctx.stark_prove(params, tr, mat, capacity=int(ctx.lib.ms_stark_proof_bound(field, params, n, 2 * w)), constants=cst)
    finally:
        ctx.close()
    assert raw == want
