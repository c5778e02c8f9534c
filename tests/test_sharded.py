"""Multi-GPU commitments (ministark_b200/sharded.py).

CPU part: the sharding plans, and the whole exchange (grouped send/recv, subtree digests, all-gather,
join) driven over `gloo` with world size 2 and 4, with the oracle standing in for the CUDA calls, so
the host logic is covered without a GPU.  GPU part: the hooked prover at world size 1 must give the
bytes of the plain prover, and with >= 2 GPUs the sharded proof must equal the single-GPU one."""
import ctypes
import hashlib
import os
import socket

import numpy as np
import pytest

from ministark_b200.sharded import ShardedCommitter, SubtreePlan, column_ranges, exchange_plan, owner_of_column

GL = 0


def test_column_ranges_cover_and_balance():
    for cols in (1, 6, 32, 33, 64):
        for world in (1, 2, 4, 8):
            r = column_ranges(cols, world)
            assert r[0][0] == 0 and r[-1][1] == cols and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
            for c in range(cols):
                g = owner_of_column(c, cols, world)
                assert r[g][0] <= c < r[g][1]


def test_exchange_plan_is_a_permutation():
    cols, world = 12, 4
    sent, received = set(), set()
    for rank in range(world):
        sends, recvs = exchange_plan(cols, world, rank)
        for _lc, c, h in sends:
            assert owner_of_column(c, cols, world) == rank and h != rank
            sent.add((c, rank, h))
        for c, g in recvs:
            assert owner_of_column(c, cols, world) == g and g != rank
            received.add((c, g, rank))
    assert sent == received and len(sent) == cols * (world - 1)


@pytest.mark.parametrize("groups,k,world,levels,left", [(1 << 10, 2, 8, 7, 1), (1 << 10, 4, 4, 4, 1), (1 << 10, 4, 2, 4, 2),
                                                       (1 << 9, 8, 8, 2, 1), (1 << 9, 8, 2, 2, 4), (16, 2, 1, 4, 1)])
def test_subtree_plan(groups, k, world, levels, left):
    p = SubtreePlan.make(groups, k, world)
    assert (p.levels_local, p.digests_per_rank) == (levels, left)


def test_subtree_plan_rejects_non_full_trees():
    with pytest.raises(ValueError):
        SubtreePlan.make(1 << 9, 4, 1)  # 2^9 leaf groups are not a power of 4 (merkle.rs:93-104)
    with pytest.raises(ValueError):
        SubtreePlan.make(24, 2, 4)


def test_committer_validates_every_shape_condition_before_any_collective():
    """the shape checks of both hooks as one pure function of the shape: all ranks raise together, before the first collective"""
    from ministark_b200.sharded import ShardedCommitter

    com = ShardedCommitter(None, 8, 2)
    com.world = 4  # the plan arithmetic only; no process group is touched
    com.validate(1 << 10, 4, 8, 4)
    for n, w, cols, blowup in [(1 << 10, 3, 8, 4),      # 3072 leaves do not divide into groups of 8 (merkle.rs:99)
                               (1 << 10, 4, 6, 4),      # leaf groups of the LDE tree != its columns
                               (24, 4, 8, 4)]:          # 12 leaf groups over 4 ranks: 3 per rank, not a power of two
        with pytest.raises(ValueError):
            com.validate(n, w, cols, blowup)
    com4 = ShardedCommitter(None, 8, 4)
    with pytest.raises(ValueError):
        com4.validate(1 << 10, 4, 8, 2)                 # 2^9 trace leaf groups: not a full 4-ary tree (merkle.rs:93-104)


# ------------------------------------------------------------------------------------------ gloo
class OracleOps:
    """Test stand-in for CudaOps: same interface, CPU tensors, compute by the oracle."""

    elem = 8

    def __init__(self, field):
        from oracle import oracle as O

        self.O, self.field = O, field

    def empty(self, *shape):
        import torch

        return torch.zeros(*shape, dtype=torch.int64)

    def empty_digests(self, n):
        import torch

        return torch.zeros(n, 8, dtype=torch.int32)

    @staticmethod
    def _view(ptr, count):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(count,))

    def lde(self, coeffs_ptr, n, ncols, blowup, shift, out):
        import torch

        coeffs = self._view(coeffs_ptr, n * ncols).reshape(ncols, n).copy()
        ev = self.O.coset_lde(self.field, coeffs, n * blowup, shift)  # [L, ncols]
        out[:ncols] = torch.from_numpy(np.ascontiguousarray(ev.T).view(np.int64))

    @staticmethod
    def _words(digest: bytes):
        return np.frombuffer(digest, dtype=">u4").astype(np.uint32).view(np.int32)

    def subtree(self, data_ptr, stride, rows, width, lpn, k, out_digests):
        import torch

        cols = [self._view(data_ptr + c * stride * 8, rows) for c in range(width)]
        flat = np.stack(cols, axis=1).reshape(-1)  # row-major flattening
        groups = flat.size // lpn
        lv = groups
        while lv > 1 and lv % k == 0:
            lv //= k
        per = flat.size // lv
        for i in range(lv):
            out_digests[i] = torch.from_numpy(self._words(self.O.merkle(flat[i * per:(i + 1) * per], lpn, k)).copy())
        return lv

    def reduce(self, digests, n, k):
        level = [digests[i].numpy().view(np.uint32).astype(">u4").tobytes() for i in range(n)]
        while len(level) > 1:
            level = [hashlib.sha256(b"".join(level[i:i + k])).digest() for i in range(0, len(level), k)]
        return level[0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, n, cols, blowup, k, w, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.synth import synth_trace

        coeffs = np.ascontiguousarray(synth_trace(GL, n, cols, seed=5).T)  # [cols, n], same on every rank
        com = ShardedCommitter(OracleOps(GL), cols, k, dist)
        lde_root = com.lde_commit(coeffs.ctypes.data, n, cols, blowup, 7)
        trace_cm = np.ascontiguousarray(synth_trace(GL, n, w, seed=9).T)  # [w, n]
        trace_root = com.trace_commit(trace_cm.ctypes.data, n, w)
        q.put((rank, lde_root, trace_root, com.stats["exchange_bytes_out"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 2), (4, 2), (2, 4), (4, 4)])
def test_sharded_commit_over_gloo_matches_oracle(world, k, oracle):
    import torch.multiprocessing as mp

    from tests.synth import synth_trace

    n, cols, blowup, w = 256, 8, 4, 8
    L = n * blowup
    coeffs = np.ascontiguousarray(synth_trace(GL, n, cols, seed=5).T)
    want_lde = oracle.merkle(oracle.coset_lde(GL, coeffs, L, 7).reshape(-1), cols, k)
    want_trace = oracle.merkle(synth_trace(GL, n, w, seed=9).reshape(-1), cols, k)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, cols, blowup, k, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _rank, lde_root, trace_root, sent in got:
        assert lde_root == want_lde
        assert trace_root == want_trace
        assert sent == (cols // world) * (world - 1) * (L // world) * 8  # L*C*s*(G-1)/G^2 per rank


# ------------------------------------------------------------------------------------------ GPU
class _NoCudaLib:
    """ms_host_register / ms_host_unregister stand-ins: the shared-memory plumbing of SharedProofBuffer is host
    code and is exercised here without a GPU (page-locking is the only CUDA call it makes)."""

    def __init__(self):
        self.registered = []

    def ms_host_register(self, h, ptr, nbytes):
        self.registered.append((ptr.value, nbytes))
        return 0

    def ms_host_unregister(self, h, ptr):
        return 0


class _NoCudaCtx:
    h = None

    def __init__(self):
        self.lib = _NoCudaLib()

    def _check(self, rc):
        assert rc == 0


def _shared_buffer_worker(rank, world, port, nbytes, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ministark_b200.sharded import SharedProofBuffer

        ctx = _NoCudaCtx()
        buf = SharedProofBuffer(ctx, nbytes, dist)
        assert buf.array.size == nbytes and ctx.lib.registered == [(buf.array.ctypes.data, nbytes)]
        # every rank writes the stripes a sharded download would give it (every world-th block of 1000 bytes)
        for blk in range(rank, nbytes // 1000, world):
            buf.array[blk * 1000:(blk + 1) * 1000] = (blk * 7 + 3) % 251
        dist.barrier()
        digest = hashlib.sha256(buf.array.tobytes()).hexdigest()  # every rank sees every rank's bytes
        buf.close()
        q.put((rank, digest))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_proof_buffer_over_gloo(world):
    """One POSIX shared-memory proof buffer per node: created by rank 0, attached (and page-locked) by the others,
    written in disjoint stripes by all ranks, identical from every rank's point of view, unlinked at close."""
    import torch.multiprocessing as mp

    nbytes = 64 * 1000
    want = np.zeros(nbytes, dtype=np.uint8)
    for blk in range(nbytes // 1000):
        want[blk * 1000:(blk + 1) * 1000] = (blk * 7 + 3) % 251
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_shared_buffer_worker, args=(r, world, port, nbytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(d == hashlib.sha256(want.tobytes()).hexdigest() for _, d in got)


@pytest.mark.gpu
@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("log_n,w,blowup,k", [(8, 4, 8, 2), (10, 8, 4, 8), (11, 4, 8, 4)])
def test_hooked_prover_world1_equals_plain(field, log_n, w, blowup, k):
    from ministark_b200 import Context
    from ministark_b200._lib import StarkParams
    from ministark_b200.sharded import stark_prove_sharded
    from tests.synth import synth_linear_matrix, synth_trace

    ctx = Context(field)
    n = 1 << log_n
    trace = synth_trace(field, n, w, seed=77 + log_n)
    m = synth_linear_matrix(field, n, w)
    params = StarkParams(40, blowup, n - 1, 2 * w, k)
    want = ctx.stark_prove(params, trace, m)
    out = np.empty(len(want) + 4096, dtype=np.uint8)
    ln = stark_prove_sharded(ctx, params, ctx.to_device(np.ascontiguousarray(trace.T)), m, out)
    assert out[:ln].tobytes() == want
    ctx.close()


def _nccl_worker(rank, world, port, log_n, w, blowup, k, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from ministark_b200 import Context
        from ministark_b200._lib import StarkParams
        from ministark_b200.sharded import stark_prove_sharded
        from tests.synth import synth_linear_matrix, synth_trace

        ctx = Context(0, rank)
        n = 1 << log_n
        trace = synth_trace(0, n, w, seed=123)
        m = synth_linear_matrix(0, n, w)
        params = StarkParams(40, blowup, n - 1, 2 * w, k)
        bound = int(ctx.lib.ms_stark_proof_bound(0, params, n, 2 * w))
        d_trace = ctx.to_device(np.ascontiguousarray(trace.T))
        if os.environ.get("MINISTARK_TEST_SHARED_DOWNLOAD") == "1":
            # one shared host buffer, every rank downloads its share of the quotient polynomials
            from ministark_b200.sharded import SharedProofBuffer

            shared = SharedProofBuffer(ctx, bound, dist)
            ln = stark_prove_sharded(ctx, params, d_trace, m, shared, dist)
            digest = hashlib.sha256(shared.array[:ln].tobytes()).hexdigest()
            shared.close()
        else:
            out = np.empty(bound, dtype=np.uint8)
            ln = stark_prove_sharded(ctx, params, d_trace, m, out, dist, proof_on_all_ranks=True)
            digest = hashlib.sha256(out[:ln].tobytes()).hexdigest()
        q.put((rank, digest))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("k", [2, 4])
def test_sharded_prover_multi_gpu_equals_single(k, exchange, monkeypatch):
    """peer: the leaf kernel reads the other ranks' LDE columns over NVLink (CUDA IPC, no exchange pass);
    nccl: grouped send/recv into a row block.  Both must give the single-GPU proof byte for byte."""
    import torch

    monkeypatch.setenv("MINISTARK_EXCHANGE", exchange)  # inherited by the spawned ranks
    # the peer runs also exercise the sharded proof download (SharedProofBuffer)
    monkeypatch.setenv("MINISTARK_TEST_SHARED_DOWNLOAD", "1" if exchange == "peer" else "0")
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    if world == 3:
        world = 2
    from ministark_b200 import Context
    from ministark_b200._lib import StarkParams
    from tests.synth import synth_linear_matrix, synth_trace

    log_n, w, blowup = 11, 4, 8  # N/2 = 4^5 trace leaf groups, L = 4^7 rows: full trees for k = 2 and 4
    n = 1 << log_n
    ctx = Context(0)
    want = ctx.stark_prove(StarkParams(40, blowup, n - 1, 2 * w, k), synth_trace(0, n, w, seed=123), synth_linear_matrix(0, n, w))
    ctx.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_nccl_worker, args=(r, world, port, log_n, w, blowup, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(h == hashlib.sha256(want).hexdigest() for _, h in got)
