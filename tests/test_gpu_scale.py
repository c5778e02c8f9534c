"""GPU parity at BASELINE sizes: whole proofs byte-for-byte against the C restatement of the
reference prover (oracle/prover.inc, or_stark_prove) run live on the box's host cores, and against the digests that
tests/golden/make_golden.py committed from that same oracle; the coset LDE at the planner's largest shapes (2^23, 2^24
rows: coset-split tiles) and widest batches (32 / 64 columns); the query phase as a stage (ms_fri_query) with forced
duplicate codeword values, 20-level paths and quotients of 2^19 coefficients."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GL, BB = 0, 1
P = {GL: 2**64 - 2**32 + 1, BB: 2013265921}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "scale_proofs.json")
THREADS = os.cpu_count() or 1


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


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


def _gpu_proof(ctx, g):
    from ministark_b200._lib import StarkParams
    from ministark_b200.synth import synth_linear_matrix, synth_trace

    n = 1 << g["log_rows"]
    tr = synth_trace(g["field"], n, g["w"], seed=g["trace_seed"])
    mat = synth_linear_matrix(g["field"], n, g["w"])
    params = StarkParams(g["security_bits"], g["blowup"], n - 1, 2 * g["w"], g["inner_children"])
    bound = int(ctx.lib.ms_stark_proof_bound(g["field"], params, n, 2 * g["w"]))
    return ctx.stark_prove(params, tr, mat, capacity=bound), tr, mat


SMALL = ["gl_2^14x8_b4", "bb_2^14x8_b4", "gl_2^16x16_b8", "bb_2^16x16_b8", "gl_2^15x8_b8_4ary", "bb_2^15x8_b8_4ary", "gl_2^16x64_b4"]


@pytest.mark.parametrize("name", SMALL)
def test_whole_proof_equals_the_c_oracle(name, ctxs, golden, oracle):
    """N = 2^14 .. 2^16, both fields, blowup 4 and 8, binary and 4-ary trees, 8 to 64 columns: proof bytes equal the C
    restatement's (src/starks.rs:59-169, src/fri.rs:53-189), and both equal the committed digest."""
    g = golden[name]
    raw, tr, mat = _gpu_proof(ctxs[g["field"]], g)
    want = oracle.stark_prove(g["field"], g["security_bits"], g["blowup"], (1 << g["log_rows"]) - 1, 2 * g["w"], tr, mat,
                              inner_children=g["inner_children"], threads=THREADS).tobytes()
    assert len(raw) == len(want) == g["proof_len"]
    assert raw == want
    assert hashlib.sha256(raw).hexdigest() == g["proof_sha256"]
    cons = oracle.derive_constrains(g["field"], tr, mat, threads=THREADS)
    assert oracle.stark_verify(g["field"], g["security_bits"], g["blowup"], (1 << g["log_rows"]) - 1, 2 * g["w"], cons, raw,
                               inner_children=g["inner_children"], strict=True) == (True, 0)


def test_config3a_proof_equals_the_c_oracle(ctxs, golden, oracle):
    """BASELINE config 3 with the binary trees StarkConfig::new builds (src/starks.rs:283-302): 2^20 rows x 16 columns,
    blowup 8 (LDE 2^23 rows), FRI to final degree; 200 MB of proof compared byte for byte with the live C oracle."""
    g = golden["config3a_gl_2^20x16_b8"]
    raw, tr, mat = _gpu_proof(ctxs[GL], g)
    assert len(raw) == g["proof_len"] and hashlib.sha256(raw).hexdigest() == g["proof_sha256"]
    want = oracle.stark_prove(GL, g["security_bits"], g["blowup"], (1 << 20) - 1, 16, tr, mat, threads=THREADS)
    assert np.array_equal(np.frombuffer(raw, dtype=np.uint8), want)


def test_headline_proof_matches_the_oracle_digest_and_the_strict_verifier(ctxs, golden, oracle):
    """BASELINE headline shape (2^22 x 32, blowup 4, 100 bits): the 940 MB proof hashes to the digest the C oracle
    produced (tests/golden/make_golden.py), and the C restatement of Stark::verify + Fri::verify (strict: Merkle paths
    enforced) accepts it."""
    g = golden["headline_gl_2^22x32_b4"]
    raw, tr, mat = _gpu_proof(ctxs[GL], g)
    assert len(raw) == g["proof_len"]
    assert hashlib.sha256(raw).hexdigest() == g["proof_sha256"]
    cons = oracle.derive_constrains(GL, tr, mat, threads=THREADS)
    assert oracle.stark_verify(GL, g["security_bits"], g["blowup"], (1 << 22) - 1, 32, cons, raw, strict=True) == (True, 0)


@pytest.mark.parametrize("field,log_n,blowup,cols", [(GL, 23, 4, 2), (GL, 24, 4, 2), (GL, 23, 8, 1), (BB, 23, 4, 2), (BB, 24, 4, 2), (BB, 24, 8, 1),
                                                     (GL, 16, 4, 32), (GL, 16, 4, 64), (BB, 16, 4, 32), (BB, 16, 8, 64), (GL, 13, 8, 64), (BB, 14, 2, 33)])
def test_coset_lde_largest_shapes_and_widest_batches(field, log_n, blowup, cols, ctxs, oracle):
    """src/starks.rs:82-91 at the sizes the planner treats specially: 2^23 / 2^24 rows (the cosets no longer fit one
    pass-1 tile and are split over tiles, ntt.cuh ntt_plan) and 32 / 64-column batches (BASELINE configs 4, 5)."""
    ctx = ctxs[field]
    n = 1 << log_n
    coeffs = rand_field(field, (cols, n), 1000 + log_n + cols)
    shift = int(rand_field(field, (1,), 9)[0]) or 5
    want = oracle.coset_lde(field, coeffs, n * blowup, shift, threads=THREADS)  # row-major [L, cols]
    got = ctx.to_host(ctx.coset_lde(ctx.to_device(coeffs), blowup, shift))  # [cols, L]
    assert np.array_equal(got.T, want)


@pytest.mark.parametrize("field", [GL, BB])
@pytest.mark.parametrize("log_domain,blowup,period", [(21, 4, 1), (21, 4, 4), (12, 2, 2), (3, 2, 1)])
def test_fri_query_stage(field, log_domain, blowup, period, ctxs, oracle):
    """ms_fri_query against the oracle on two consecutive committed rounds: points, first-match value search on a
    codeword with forced duplicates (a polynomial in x^period takes every value `period` times, SURVEY hard part 9),
    leaf neighbours, sibling pairs of a 20-level tree, and the long-division quotients (src/fri.rs:148-171,
    src/merkle.rs:216-288)."""
    ctx = ctxs[field]
    D, p = ctx.D, P[field]
    nd = 1 << log_domain
    npad = nd // blowup
    plen = npad - (npad // 8 if npad >= 16 else 0)
    if period > 1:
        plen -= plen % period
    poly = rand_field(field, (plen, D), 5 + log_domain)
    if period > 1:
        mask = (np.arange(plen) % period) != 0
        poly[mask] = 0
        while plen > 0 and not poly[plen - 1].any():
            plen -= 1
        poly = poly[:plen]
    # next round's polynomial: any polynomial of half the size (the query phase only reads its codeword)
    nxt = rand_field(field, (max(npad // 2, 1), D), 6)
    planes = np.zeros((D, npad), dtype=np.uint64)
    planes[:, :plen] = poly.T
    nplanes = np.zeros((D, max(npad // 2, 1)), dtype=np.uint64)
    nplanes[:, : nxt.shape[0]] = nxt.T
    d_poly, d_next = ctx.to_device(planes), ctx.to_device(nplanes)
    cw, nodes, _root = ctx.fri_commit(d_poly, nd, blowup)
    ncw, _nn, _nr = ctx.fri_commit(d_next, nd // 2, nd // 2 // nplanes.shape[1])
    rng = np.random.default_rng(3)
    betas = [int(b) for b in rng.integers(0, 2**63, size=6, dtype=np.uint64)] + [0, nd, nd + 1, nd - 1, 1]
    got = ctx.fri_query(d_poly, plen, cw, nodes, ncw, betas)
    # ---- oracle side
    o_cw = oracle.fri_codeword(field, poly, nd)  # [nd, D]
    o_ncw = oracle.fri_codeword(field, nxt, nd // 2)
    o_root, o_nodes = oracle.merkle(o_cw, 2, 2, deg=D, want_nodes=True)
    assert np.array_equal(ctx.nodes_to_bytes(nodes), o_nodes)
    g_prev, g_next = oracle.root_of_unity(field, log_domain), oracle.root_of_unity(field, log_domain - 1)
    path_len = log_domain - 1
    for k, beta in enumerate(betas):
        b = beta % nd if beta > nd else beta  # fri.rs:144 (strict >)
        x1, x2, x3 = pow(g_prev, b, p), pow(g_prev, nd // 2 + b, p), pow(g_next, b, p)
        y1 = oracle.eval_ext_at_ext(field, poly, [x1] + [0] * (D - 1))
        y2 = oracle.eval_ext_at_ext(field, poly, [x2] + [0] * (D - 1))
        y3 = oracle.eval_ext_at_ext(field, nxt, [x3] + [0] * (D - 1))
        pts = got["points"][k]
        assert [int(v) for v in pts[0]] == [x1] + [0] * (D - 1) and [int(v) for v in pts[2]] == [x2] + [0] * (D - 1)
        assert [int(v) for v in pts[4]] == [x3] + [0] * (D - 1)
        assert np.array_equal(pts[1], y1) and np.array_equal(pts[3], y2) and np.array_equal(pts[5], y3)
        for which, y in enumerate((y1, y2)):
            first = int(np.argmax((o_cw == y).all(axis=1)))  # merkle.rs:216-225: first index holding the value
            m = 2 * k + which
            assert int(got["found"][m]) == first
            if period > 1 and nd > 8:
                assert first < nd // period  # the duplicates are real: the search lands in the first period
            assert np.array_equal(got["neigh"][m], o_cw[first - first % 2: first - first % 2 + 2])
            off, lv, node = 0, nd // 2, first // 2
            for l in range(path_len):
                pair = node - node % 2
                assert np.array_equal(got["paths"][m, l].reshape(-1), o_nodes[off + pair: off + pair + 2].reshape(-1)), (k, which, l)
                off += lv
                lv //= 2
                node //= 2
        if plen >= 3:
            q = oracle.fri_query_quotient(field, poly, x1, x2, y1, y2)
            want_q = np.zeros((plen - 2, D), dtype=np.uint64)
            want_q[: q.shape[0]] = q
            assert np.array_equal(got["quot"][k], want_q)
