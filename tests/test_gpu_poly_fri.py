"""GPU parity, coefficient-space and FRI stages (a3, a6, a7, a8-a10 and the query-phase quotient)
against the C oracle with injected challenges (independent of the transcript)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GL, BB = 0, 1
P = {GL: 2**64 - 2**32 + 1, BB: 2013265921}


def rand_field(field, shape, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 2**63, size=shape, dtype=np.uint64) * 2 + rng.integers(0, 2, size=shape, dtype=np.uint64)) % np.uint64(P[field])


@pytest.fixture(scope="module")
def ctxs():
    from ministark_b200 import Context

    c = {GL: Context(GL), BB: Context(BB)}
    yield c
    for v in c.values():
        v.close()


@pytest.mark.parametrize("field", [GL, BB])
@pytest.mark.parametrize("n,w", [(16, 3), (1 << 12, 8), (1 << 15, 5)])
def test_linear_constraints_and_mix(field, n, w, ctxs, oracle):
    from tests.synth import synth_linear_matrix

    ctx = ctxs[field]
    p = P[field]
    coeffs = rand_field(field, (w, n), n + w)
    m = synth_linear_matrix(field, n, w)
    got = ctx.to_host(ctx.linear_constraints(ctx.to_device(coeffs), m))
    for t in range(w):  # every coefficient, exact big-int arithmetic
        want = (int(m[t, t]) * coeffs[t].astype(object) + int(m[t, (t + 1) % w]) * coeffs[(t + 1) % w].astype(object)) % p
        assert (got[t].astype(object) == want).all()
    # the same columns through the C oracle's derive_constrains-style accumulation (dense matrix, incl. zero entries)
    dense = rand_field(field, (3, w), 17)
    dense[1, :] = 0
    got_d = ctx.to_host(ctx.linear_constraints(ctx.to_device(coeffs), dense))
    for t in range(3):
        want = sum(int(dense[t, j]) * coeffs[j].astype(object) for j in range(w)) % p
        assert (got_d[t].astype(object) == want).all()
    # e2e_goldilocks.rs:57-59 style row with three non-zeros
    m2 = np.zeros((1, w), dtype=np.uint64)
    m2[0, :3] = [p - 1, p - 1, 1]
    got2 = ctx.to_host(ctx.linear_constraints(ctx.to_device(coeffs), m2))[0]
    want2 = (coeffs[2].astype(object) - coeffs[0].astype(object) - coeffs[1].astype(object)) % p
    assert (got2.astype(object) == want2).all()
    allc = np.concatenate([coeffs, got], axis=0)
    r = int(rand_field(field, (1,), 3)[0])
    assert (ctx.to_host(ctx.mix(ctx.to_device(allc), r)) == oracle.mix(field, allc, r)).all()


@pytest.mark.parametrize("field", [GL, BB])
@pytest.mark.parametrize("n,cols,q", [(1, 2, 1), (8, 7, 1), (2048, 3, 2), (2049 * 4, 2, 3), (1 << 16, 5, 3)])
def test_deep_open(field, n, cols, q, ctxs, oracle):
    ctx = ctxs[field]
    coeffs = rand_field(field, (cols, n), n)
    z = rand_field(field, (q, ctx.D), 5)
    got = ctx.deep_open(ctx.to_device(coeffs), z)
    for qi in range(q):
        for c in range(cols):
            assert (got[qi, c] == oracle.eval_base_at_ext(field, coeffs[c], z[qi])).all()


@pytest.mark.parametrize("field", [GL, BB])
@pytest.mark.parametrize("npad,blowup,ncoef", [(1, 2, 0), (1, 2, 1), (4, 2, 3), (16, 4, 16), (1 << 10, 8, 1000), (1 << 13, 4, (1 << 13) - 1)])
def test_fri_commit(field, npad, blowup, ncoef, ctxs, oracle):
    ctx = ctxs[field]
    D = ctx.D
    poly = np.zeros((npad, D), dtype=np.uint64)
    poly[:ncoef] = rand_field(field, (ncoef, D), npad + 1)
    domain = npad * blowup
    cw, nodes, root = ctx.fri_commit(ctx.to_device(np.ascontiguousarray(poly.T)), domain, blowup)
    want_cw = oracle.fri_codeword(field, poly, domain)
    assert (ctx.to_host(cw).T == want_cw).all()
    want_root, want_nodes = oracle.merkle(want_cw, 2, 2, deg=D, want_nodes=True)
    assert root == want_root
    assert (ctx.nodes_to_bytes(nodes) == want_nodes).all()


@pytest.mark.parametrize("field", [GL, BB])
@pytest.mark.parametrize("npad,ncoef", [(1, 1), (2, 2), (4, 3), (8, 8), (64, 63), (4096, 4095), (8192, 8191), (1 << 15, (1 << 15) - 1), (1 << 17, 1 << 17)])
def test_fri_deep_and_fold(field, npad, ncoef, ctxs, oracle):
    ctx = ctxs[field]
    D = ctx.D
    poly = np.zeros((npad, D), dtype=np.uint64)
    poly[:ncoef] = rand_field(field, (ncoef, D), npad + 7)
    z, alpha = rand_field(field, (D,), 1), rand_field(field, (D,), 2)
    want_d, want_next = oracle.fri_fold(field, poly[:ncoef], z, alpha)
    planes = ctx.to_device(np.ascontiguousarray(poly.T))
    d = ctx.fri_deep_coeffs(planes, z)
    assert (d == want_d).all()
    nxt = ctx.to_host(ctx.fri_fold(planes, z, alpha, d)).T  # [max(npad/2,1), D], zero padded
    assert (nxt[: want_next.shape[0]] == want_next).all()
    assert not nxt[want_next.shape[0]:].any()


def test_fri_fold_is_exact_division(ctxs, oracle):
    """size-independent property: next(x) * (x - z) + d(alpha) == folded(x) at a random point."""
    ctx = ctxs[GL]
    npad = 1 << 18
    poly = rand_field(GL, (npad, 2), 99)
    z, alpha, t = rand_field(GL, (2,), 1), rand_field(GL, (2,), 2), rand_field(GL, (2,), 3)
    planes = ctx.to_device(np.ascontiguousarray(poly.T))
    d = ctx.fri_deep_coeffs(planes, z)
    nxt = ctx.to_host(ctx.fri_fold(planes, z, alpha, d)).T
    ev = lambda coef, x: oracle.eval_ext_at_ext(GL, np.ascontiguousarray(coef), x)
    mul = lambda a, b: oracle.ext_mul(GL, a, b)
    add = lambda a, b: (a.astype(object) + b.astype(object)) % P[GL]
    sub = lambda a, b: (a.astype(object) - b.astype(object)) % P[GL]
    u64 = lambda a: np.array([int(v) for v in a], dtype=np.uint64)
    folded_t = add(ev(poly[0::2], t), mul(alpha, ev(poly[1::2], t)))
    lhs = add(mul(ev(nxt, t), u64(sub(t, z))), add(d[0], mul(d[1], alpha)))
    assert (u64(lhs) == u64(folded_t)).all()
