"""The multi-GPU prover inside libministark.so (csrc/comm.cuh, csrc/prover.cuh).

  * "virtual ranks": G contexts of ONE process bound with ms_comm_init_local and driven by G host threads -- on the
    1-GPU test box all of them share the one B200, which exercises every sharded code path (row-sharded trees, column-
    sharded iNTT / constraints / LDE / mixing / openings, peer-read leaf hashing and path gathering, sharded download);
  * NCCL: one process per GPU through the C ABI only (ctypes; no torch.distributed), id exchanged through a file
    (needs >= 2 GPUs, skipped otherwise).
The proof must be byte-identical to the single-GPU proof -- and therefore to the C oracle's (tests/golden) -- for every
number of ranks, every shard mask and both download modes."""
import ctypes as C
import hashlib
import json
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GL, BB = 0, 1
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "scale_proofs.json")


def _inputs(field, log_n, w, seed=0x5EED000000000000, t=None):
    from ministark_b200.synth import synth_linear_matrix, synth_trace

    n = 1 << log_n
    tr = synth_trace(field, n, w, seed=seed)
    m = synth_linear_matrix(field, n, w)
    if t is not None:
        m = m[:t]
    return tr, m


def _prove_single(field, params, tr, m):
    from ministark_b200 import Context

    ctx = Context(field)
    try:
        bound = int(ctx.lib.ms_stark_proof_bound(field, params, tr.shape[0], tr.shape[1] + m.shape[0]))
        return ctx.stark_prove(params, tr, m, capacity=bound)
    finally:
        ctx.close()


def _prove_virtual(field, params, tr, m, world, shared, mask=15):
    """world contexts on device 0, one thread each; returns rank 0's proof bytes"""
    import torch

    from ministark_b200 import Context

    ctxs = [Context(field, 0) for _ in range(world)]
    try:
        Context.comm_init_local(ctxs)
        assert ctxs[1].comm_info() == (1, world, "local")
        for c in ctxs:
            c.set_shard_mask(mask)
        n, w = tr.shape
        bound = int(ctxs[0].lib.ms_stark_proof_bound(field, params, n, w + m.shape[0]))
        trace_cm = ctxs[0].to_device(np.ascontiguousarray(tr.T))
        torch.cuda.synchronize()
        bufs = [np.zeros(bound, dtype=np.uint8)] * world if shared else [np.zeros(bound, dtype=np.uint8)] + [None] * (world - 1)
        lens, errs = [0] * world, [None] * world

        def run(g):
            try:
                for _ in range(2):  # twice: the arena and its reuse across proofs
                    lens[g] = ctxs[g].stark_prove_multi(params, trace_cm, m, bufs[g], shared=shared)
            except Exception as e:  # noqa: BLE001
                errs[g] = e

        th = [threading.Thread(target=run, args=(g,)) for g in range(world)]
        for x in th:
            x.start()
        for x in th:
            x.join(timeout=600)
        assert not any(x.is_alive() for x in th), "a rank hangs"
        for e in errs:
            if e is not None:
                raise e
        assert len(set(lens)) == 1
        return bufs[0][: lens[0]].tobytes()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("name,world,shared", [("gl_2^14x8_b4", 2, False), ("gl_2^14x8_b4", 8, True), ("bb_2^14x8_b4", 4, True),
                                               ("gl_2^16x16_b8", 4, True), ("bb_2^16x16_b8", 8, False), ("gl_2^15x8_b8_4ary", 4, True),
                                               ("bb_2^15x8_b8_4ary", 2, True), ("gl_2^16x64_b4", 8, True)])
def test_virtual_ranks_reproduce_the_oracle_proof(name, world, shared):
    """golden shapes (C oracle digests) proved by G virtual ranks: same bytes as the oracle's"""
    from ministark_b200._lib import StarkParams

    with open(GOLDEN) as fh:
        g = json.load(fh)[name]
    tr, m = _inputs(g["field"], g["log_rows"], g["w"], seed=g["trace_seed"])
    params = StarkParams(g["security_bits"], g["blowup"], (1 << g["log_rows"]) - 1, 2 * g["w"], g["inner_children"])
    raw = _prove_virtual(g["field"], params, tr, m, world, shared)
    assert len(raw) == g["proof_len"] and hashlib.sha256(raw).hexdigest() == g["proof_sha256"]


@pytest.mark.parametrize("field,log_n,w,t,blowup,lpn,k,world,mask", [
    (GL, 12, 3, 3, 4, 6, 2, 2, 15),    # W = T = 3 over 2 ranks: uneven column shares
    (GL, 12, 2, 2, 2, 4, 2, 4, 15),    # fewer columns and constraint rows than ranks (some ranks own none)
    (BB, 13, 3, 3, 8, 6, 2, 8, 15),    # fewer columns than ranks
    (GL, 12, 4, 4, 4, 4, 2, 4, 15),    # leaf groups != columns: the LDE path falls back to replicas, trees still shard where they can
    (GL, 14, 4, 4, 4, 8, 2, 4, 1),     # one stage sharded at a time
    (GL, 14, 4, 4, 4, 8, 2, 4, 2),
    (GL, 14, 4, 4, 4, 8, 2, 4, 4),
    (BB, 14, 4, 4, 4, 8, 2, 2, 8),
    (GL, 14, 4, 4, 4, 8, 2, 3, 15),    # 3 ranks: trees do not split, columns do not either (2^16 rows / 3): all replicas
    (GL, 4, 2, 2, 2, 4, 2, 2, 15),     # tiny: every tree smaller than the rank count thresholds
])
def test_virtual_ranks_equal_single_gpu(field, log_n, w, t, blowup, lpn, k, world, mask):
    from ministark_b200._lib import StarkParams

    tr, m = _inputs(field, log_n, w, t=t)
    params = StarkParams(40, blowup, (1 << log_n) - 1, lpn, k)
    want = _prove_single(field, params, tr, m)
    for shared in (False, True):
        got = _prove_virtual(field, params, tr, m, world, shared, mask)
        assert got == want, (shared,)


def _nccl_rank(rank, world, idfile, field, log_n, w, q):
    """one process per GPU, C ABI only"""
    try:
        import time

        import torch

        from ministark_b200 import Context
        from ministark_b200._lib import StarkParams

        torch.cuda.set_device(rank)
        ctx = Context(field, rank)
        if rank == 0:
            uid = Context.comm_unique_id()
            with open(idfile + ".tmp", "wb") as fh:
                fh.write(uid)
            os.replace(idfile + ".tmp", idfile)
        else:
            for _ in range(600):
                if os.path.exists(idfile):
                    break
                time.sleep(0.05)
            with open(idfile, "rb") as fh:
                uid = fh.read()
        ctx.comm_init_nccl(uid, rank, world)
        tr, m = _inputs(field, log_n, w)
        n = 1 << log_n
        params = StarkParams(60, 4, n - 1, 2 * w, 2)
        bound = int(ctx.lib.ms_stark_proof_bound(field, params, n, 2 * w))
        trace_cm = ctx.to_device(np.ascontiguousarray(tr.T))
        buf = np.zeros(bound, dtype=np.uint8) if rank == 0 else None
        for _ in range(2):
            plen = ctx.stark_prove_multi(params, trace_cm, m, buf)
        digest = hashlib.sha256(buf[:plen].tobytes()).hexdigest() if rank == 0 else None
        info = ctx.comm_info()
        ctx.comm_destroy()
        ctx.close()
        q.put((rank, "ok", digest, info))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "error", repr(e), None))


@pytest.mark.parametrize("field,log_n,w", [(GL, 16, 8), (BB, 15, 4)])
def test_nccl_ranks_through_the_c_abi_only(field, log_n, w, tmp_path):
    """world-2 proof over NCCL + CUDA IPC with nothing but ctypes calls (no torch.distributed): ms_comm_unique_id ->
    file -> ms_comm_init_nccl -> ms_stark_prove_multi; equals the single-GPU proof."""
    import multiprocessing as mp

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from ministark_b200._lib import StarkParams

    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    idfile = str(tmp_path / "nccl_id")
    procs = [mpc.Process(target=_nccl_rank, args=(r, world, idfile, field, log_n, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    tr, m = _inputs(field, log_n, w)
    want = _prove_single(field, StarkParams(60, 4, (1 << log_n) - 1, 2 * w, 2), tr, m)
    digest = [r[2] for r in res if r[0] == 0][0]
    assert digest == hashlib.sha256(want).hexdigest()
    assert sorted(r[3] for r in res) == [(0, 2, "nccl"), (1, 2, "nccl")]


