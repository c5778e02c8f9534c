"""Multi-GPU commitments: one prover replica per GPU (one process per GPU, torch.distributed over
NCCL), with the two stages that shard -- SURVEY.md 8e -- split across the ranks:

  trace tree (src/starks.rs:70-72)   every rank holds the trace; rank h hashes the leaf groups of rows
                                     [h N/G, (h+1) N/G) and climbs its subtree; the G (or G*r) digests
                                     are all-gathered and joined on every rank.
  LDE + tree (src/starks.rs:82-94)   rank g extends columns [g C/G, (g+1) C/G) (columns are independent),
                                     then the column-to-row exchange: for every local column the row
                                     range of rank h goes to rank h (grouped NCCL send/recv, written
                                     straight into the receiver's [C][L/G] column-major block, no
                                     packing pass); rank h hashes its L/G rows, climbs, all-gather, join.

Everything else in Stark::prove (transcript, mixing, openings, FRI: a single polynomial, 8e "replicas
only") runs identically on every rank, so all ranks derive the same challenges without a broadcast
and rank 0's proof is the proof.  The exchange is the only bulk collective: L*C*s*(G-1)/G^2 bytes out
per GPU.

Two ways to do the exchange (MINISTARK_EXCHANGE = "peer" | "nccl", default "peer" on GPUs):
  peer   no exchange pass at all: every rank's LDE columns live in a CUDA-IPC buffer the other ranks
         have opened, and the leaf-hash kernel of rank h loads rows [h L/G, (h+1) L/G) of every column
         straight from its owner over NVLink while it hashes (ms_merkle_subtree_gather): the transfer
         is hidden under the SHA-256 arithmetic, and the [C][L/G] receive block is never materialised.
         Two tiny all-reduces order the ranks (all LDEs done before anyone reads; all reads done before
         anyone overwrites its columns -- the second one is the digest all-gather that is needed anyway).
  nccl   grouped ncclSend/ncclRecv into a [C][L/G] block, then a local ms_merkle_subtree (the baseline,
         and what the CPU gloo tests exercise).

The compute calls go through a small `ops` object: CudaOps (the C ABI, the product path) or, in the
CPU-only gloo tests, an oracle-backed stand-in defined under tests/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


# ------------------------------------------------------------------------------------------ plans
def column_ranges(cols: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous column shards, as even as possible (first cols % world ranks get one more)."""
    base, extra = divmod(cols, world)
    out, start = [], 0
    for g in range(world):
        n = base + (1 if g < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def owner_of_column(col: int, cols: int, world: int) -> int:
    for g, (a, b) in enumerate(column_ranges(cols, world)):
        if a <= col < b:
            return g
    raise IndexError(col)


@dataclass
class SubtreePlan:
    """How a tree with `groups` leaf groups and arity k splits over `world` ranks."""
    groups: int
    k: int
    world: int
    groups_per_rank: int
    levels_local: int      # levels each rank climbs above its leaf digests
    digests_per_rank: int  # what is left per rank (< k); world * digests_per_rank digests are joined

    @staticmethod
    def make(groups: int, k: int, world: int) -> "SubtreePlan":
        if groups % world or groups < world:
            raise ValueError(f"{groups} leaf groups do not split over {world} ranks")
        per = groups // world
        if per & (per - 1):
            raise ValueError("leaf groups per rank must be a power of two")
        lv, levels = per, 0
        while lv > 1 and lv % k == 0:
            lv //= k
            levels += 1
        total = world * lv
        t = total
        while t > 1:
            if t % k:
                raise ValueError(f"Tree is not full! {groups} leaf groups, inner_children {k} (merkle.rs:93-104)")
            t //= k
        return SubtreePlan(groups, k, world, per, levels, lv)


def exchange_plan(cols: int, world: int, rank: int):
    """(sends, recvs) of the column-to-row exchange for `rank`: sends = [(local_col, global_col, peer)],
    recvs = [(global_col, peer)]; every (column, peer != rank) pair appears exactly once on each side."""
    ranges = column_ranges(cols, world)
    a, b = ranges[rank]
    sends = [(c - a, c, h) for h in range(world) if h != rank for c in range(a, b)]
    recvs = [(c, g) for g in range(world) if g != rank for c in range(*ranges[g])]
    return sends, recvs


# ------------------------------------------------------------------------------------------ compute backends
class CudaOps:
    """The product path: libministark.so through the C ABI on this rank's GPU."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib
        self.elem = 8 if ctx.field == 0 else 4
        self._bufs = {}

    def empty(self, *shape):
        """exchange buffers are reused across proofs (a 2^24 x 32 shard is 4 GiB: no per-proof cudaMalloc)"""
        t = self._bufs.get(shape)
        if t is None:
            t = self._bufs[shape] = self.ctx.empty(*shape)
        return t

    def empty_digests(self, n):
        import torch

        return torch.empty(n, 8, dtype=torch.int32, device=f"cuda:{self.ctx.device}")

    def lde(self, coeffs_ptr: int, n: int, ncols: int, blowup: int, shift: int, out):
        self.ctx._check(self.lib.ms_coset_lde(self.ctx.h, C.c_void_p(coeffs_ptr), n, n, ncols, blowup, shift,
                                              C.c_void_p(out.data_ptr()), out.stride(0)))

    def subtree(self, data_ptr: int, stride: int, rows: int, width: int, lpn: int, k: int, out_digests) -> int:
        n_out = C.c_uint64(0)
        self.ctx._check(self.lib.ms_merkle_subtree(self.ctx.h, C.c_void_p(data_ptr), stride, rows, width, 1, lpn, k,
                                                   C.c_void_p(out_digests.data_ptr()), C.byref(n_out)))
        return int(n_out.value)

    def reduce(self, digests, n: int, k: int) -> bytes:
        root = (C.c_uint8 * 32)()
        self.ctx._check(self.lib.ms_merkle_reduce(self.ctx.h, C.c_void_p(digests.data_ptr()), n, k, root))
        return bytes(root)

    # ---- peer memory (CUDA IPC) for the exchange-free LDE tree
    def lde_raw(self, coeffs_ptr: int, n: int, ncols: int, blowup: int, shift: int, out_ptr: int, out_stride: int):
        self.ctx._check(self.lib.ms_coset_lde(self.ctx.h, C.c_void_p(coeffs_ptr), n, n, ncols, blowup, shift,
                                              C.c_void_p(out_ptr), out_stride))

    def subtree_gather(self, plane_ptrs: Sequence[int], rows: int, width: int, lpn: int, k: int, out_digests) -> int:
        tab = (C.c_void_p * len(plane_ptrs))(*plane_ptrs)
        n_out = C.c_uint64(0)
        self.ctx._check(self.lib.ms_merkle_subtree_gather(self.ctx.h, tab, rows, width, 1, lpn, k,
                                                          C.c_void_p(out_digests.data_ptr()), C.byref(n_out)))
        return int(n_out.value)

    def peer_buffers(self, key, nbytes: int, dist, group):
        """This rank's IPC-exported buffer of `nbytes` plus every rank's view of every other rank's buffer:
        returns [ptr_of_rank_0, ..., ptr_of_rank_{G-1}] (own entry = the local pointer).  Cached per key."""
        import torch

        cache = self.__dict__.setdefault("_peer", {})
        if key in cache:
            return cache[key][0]
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = C.c_void_p()
        self.ctx._check(self.lib.ms_peer_alloc(self.ctx.h, nbytes, C.byref(mine)))
        handle = (C.c_uint8 * 64)()
        self.ctx._check(self.lib.ms_peer_export(self.ctx.h, mine, handle))
        dev = f"cuda:{self.ctx.device}"
        local = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, local, group=group)
        allh = allh.cpu().numpy()
        ptrs = []
        for g in range(world):
            if g == rank:
                ptrs.append(int(mine.value))
                continue
            h = (C.c_uint8 * 64)(*allh[g * 64:(g + 1) * 64].tolist())
            p = C.c_void_p()
            self.ctx._check(self.lib.ms_peer_open(self.ctx.h, h, C.byref(p)))
            ptrs.append(int(p.value))
        cache[key] = (ptrs, rank)
        return ptrs

    def close_peers(self, dist=None, group=None):
        """Unmap the peers' buffers and free the own ones (every rank must call it: a barrier first makes sure
        nobody still reads)."""
        cache = self.__dict__.pop("_peer", {})
        if not cache:
            return
        if dist is not None:
            dist.barrier(group=group)
        for ptrs, rank in cache.values():
            for g, p in enumerate(ptrs):
                if g != rank:
                    self.lib.ms_peer_close(self.ctx.h, C.c_void_p(p))
        if dist is not None:
            dist.barrier(group=group)
        for ptrs, rank in cache.values():
            self.lib.ms_peer_free(self.ctx.h, C.c_void_p(ptrs[rank]))



# ------------------------------------------------------------------------------------------ the committer
class ShardedCommitter:
    """Implements ms_commit_hooks for one rank.  `dist` is torch.distributed (initialised) or None for a
    single process; `ops` the compute backend."""

    def __init__(self, ops, leafs_per_node: int, inner_children: int = 2, dist=None, group=None):
        self.ops = ops
        self.lpn = leafs_per_node
        self.k = inner_children
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.stats = {"exchange_bytes_out": 0, "lde_cols": 0}
        self._keep = None
        import os

        want = os.environ.get("MINISTARK_EXCHANGE", "peer")
        self.exchange = "peer" if (want == "peer" and self.world > 1 and hasattr(ops, "peer_buffers")) else "nccl"

    # ---- shared: gather every rank's digests and join them
    def _join(self, mine, n_mine: int) -> bytes:
        import torch

        if self.world == 1:
            return self.ops.reduce(mine, n_mine, self.k)
        allg = torch.empty((self.world * n_mine, 8), dtype=mine.dtype, device=mine.device)
        self.dist.all_gather_into_tensor(allg, mine[:n_mine].contiguous(), group=self.group)
        return self.ops.reduce(allg, self.world * n_mine, self.k)

    def validate(self, n: int, w: int, cols: int, blowup: int) -> None:
        """Every shape condition of the two hooks, checked up front: a pure function of the shape, so all ranks raise
        together BEFORE the first collective (a rank failing alone inside a hook would leave its peers waiting in one)."""
        if n * w % self.lpn:
            raise ValueError("leaf count not divisible by leafs_per_node (merkle.rs:99)")
        plan = SubtreePlan.make(n * w // self.lpn, self.k, self.world)
        if (plan.groups_per_rank * self.lpn) % w:
            raise ValueError("a rank's leaf groups must cover whole rows")
        if self.lpn != cols:
            raise ValueError("the LDE tree hashes one row per leaf group (leafs_per_node == columns)")
        SubtreePlan.make(n * blowup, self.k, self.world)

    # ---- a1: the trace tree, row ranges, no bulk exchange (every rank already holds the trace)
    def trace_commit(self, trace_ptr: int, n: int, w: int) -> bytes:
        groups = n * w // self.lpn
        if n * w % self.lpn:
            raise ValueError("leaf count not divisible by leafs_per_node (merkle.rs:99)")
        plan = SubtreePlan.make(groups, self.k, self.world)
        elems = plan.groups_per_rank * self.lpn
        if elems % w:
            raise ValueError("a rank's leaf groups must cover whole rows")
        rows = elems // w
        digests = self.ops.empty_digests(max(plan.digests_per_rank, 1))
        got = self.ops.subtree(trace_ptr + self.rank * rows * self.ops.elem, n, rows, w, self.lpn, self.k, digests)
        assert got == plan.digests_per_rank
        return self._join(digests, got)

    # ---- a4 + a5: column-sharded LDE, column-to-row exchange, row-sharded tree
    def _mark(self):
        """device-time marker on the current stream (GPU backend only)"""
        if not getattr(self.ops, "ctx", None):
            return None
        import torch

        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def lde_commit(self, coeffs_ptr: int, n: int, cols: int, blowup: int, shift: int) -> bytes:
        L = n * blowup
        if self.lpn != cols:
            raise ValueError("the LDE tree hashes one row per leaf group (leafs_per_node == columns)")
        plan = SubtreePlan.make(L, self.k, self.world)
        rows = plan.groups_per_rank
        a, b = column_ranges(cols, self.world)[self.rank]
        if self.exchange == "peer":
            return self._lde_commit_peer(coeffs_ptr, n, cols, blowup, shift, plan)
        t0 = self._mark()
        mine = self.ops.empty(max(b - a, 1), L)
        if b > a:
            self.ops.lde(coeffs_ptr + a * n * self.ops.elem, n, b - a, blowup, shift, mine)
        self.stats["lde_cols"] = b - a
        t1 = self._mark()
        if self.world == 1:
            block = mine
        else:
            block = self.ops.empty(cols, rows)
            sends, recvs = exchange_plan(cols, self.world, self.rank)
            d = self.dist
            reqs = [d.P2POp(d.isend, mine[lc, h * rows:(h + 1) * rows], self._peer(h), self.group) for lc, _, h in sends]
            reqs += [d.P2POp(d.irecv, block[c], self._peer(g), self.group) for c, g in recvs]
            work = d.batch_isend_irecv(reqs) if reqs else []
            block[a:b].copy_(mine[:, self.rank * rows:(self.rank + 1) * rows])  # own rows: local copy
            for wk in work:
                wk.wait()
            self.stats["exchange_bytes_out"] = len(sends) * rows * self.ops.elem
        self.stats["exchange"] = "nccl-sendrecv"
        t2 = self._mark()
        digests = self.ops.empty_digests(max(plan.digests_per_rank, 1))
        got = self.ops.subtree(block.data_ptr(), block.stride(0), rows, cols, self.lpn, self.k, digests)
        assert got == plan.digests_per_rank
        t3 = self._mark()
        root = self._join(digests, got)
        if t0 is not None:
            t3.synchronize()
            self.stats.update(lde_ms=t0.elapsed_time(t1), exchange_ms=t1.elapsed_time(t2), tree_ms=t2.elapsed_time(t3))
        self._keep = (mine, block)  # alive until the stream has consumed them
        return root

    def _lde_commit_peer(self, coeffs_ptr: int, n: int, cols: int, blowup: int, shift: int, plan: SubtreePlan) -> bytes:
        """a4 + a5 without an exchange pass: the leaf kernel reads the peers' columns over NVLink."""
        import torch

        ops, d = self.ops, self.dist
        L, rows, elem = n * blowup, plan.groups_per_rank, self.ops.elem
        ranges = column_ranges(cols, self.world)
        a, b = ranges[self.rank]
        most = max(hi - lo for lo, hi in ranges)
        bases = ops.peer_buffers(("lde", most, L), most * L * elem, d, self.group)
        flag = ops.__dict__.setdefault("_flag", torch.zeros(1, dtype=torch.int32, device=f"cuda:{ops.ctx.device}"))
        t0 = self._mark()
        if b > a:
            ops.lde_raw(coeffs_ptr + a * n * elem, n, b - a, blowup, shift, bases[self.rank], L)
        self.stats["lde_cols"] = b - a
        t1 = self._mark()
        d.all_reduce(flag, group=self.group)  # every rank's columns are complete (stream-ordered)
        t2 = self._mark()
        planes = [bases[g] + ((c - lo) * L + self.rank * rows) * elem for g, (lo, hi) in enumerate(ranges) for c in range(lo, hi)]
        digests = ops.empty_digests(max(plan.digests_per_rank, 1))
        got = ops.subtree_gather(planes, rows, cols, self.lpn, self.k, digests)
        assert got == plan.digests_per_rank
        t3 = self._mark()
        root = self._join(digests, got)  # the all-gather also tells every rank that its columns were read
        self.stats["exchange_bytes_out"] = (cols - (b - a)) * rows * elem  # pulled over NVLink by the leaf kernel
        self.stats["exchange"] = "peer-read"
        if t0 is not None:
            t3.synchronize()
            self.stats.update(lde_ms=t0.elapsed_time(t1), exchange_ms=t1.elapsed_time(t2), tree_ms=t2.elapsed_time(t3))
        return root

    def _peer(self, group_rank: int) -> int:
        if self.group is None or self.dist is None:
            return group_rank
        return self.dist.get_global_rank(self.group, group_rank)

    # ---- ctypes glue
    def hooks(self, replica_only: bool = False, download: Tuple[int, int] = (0, 0)) -> "_lib.CommitHooks":
        def _tc(_user, d_trace, n, w, root32):
            try:
                root = self.trace_commit(int(d_trace), int(n), int(w))
            except Exception as e:  # never unwind across the C boundary
                self.error = e
                return 1
            C.memmove(root32, root, 32)
            return 0

        def _lc(_user, d_coeffs, n, cols, blowup, shift, root32):
            try:
                root = self.lde_commit(int(d_coeffs), int(n), int(cols), int(blowup), int(shift))
            except Exception as e:
                self.error = e
                return 1
            C.memmove(root32, root, 32)
            return 0

        self.error: Optional[Exception] = None
        self._cb = (_lib.TRACE_COMMIT_FN(_tc), _lib.LDE_COMMIT_FN(_lc))
        return _lib.CommitHooks(None, self._cb[0], self._cb[1], 1 if replica_only else 0, download[0], download[1])


def _touch_interleaved(buf, nbytes: int) -> None:
    """first-touch the buffer under MPOL_INTERLEAVE so that its pages alternate between the host's NUMA nodes (all ranks'
    downloads then spread over every memory controller instead of landing on rank 0's node)"""
    import glob

    nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*")) or 1
    libc = C.CDLL(None, use_errno=True)
    mask = C.c_ulong((1 << nodes) - 1)
    MPOL_DEFAULT, MPOL_INTERLEAVE, SYS_set_mempolicy = 0, 3, 238  # x86_64
    ok = nodes > 1 and libc.syscall(SYS_set_mempolicy, MPOL_INTERLEAVE, C.byref(mask), C.c_ulong(nodes + 1)) == 0
    a = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
    a[::4096] = 0
    if ok:
        libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, C.c_ulong(0))


class SharedProofBuffer:
    """One host buffer for the proof, visible to every rank of a node (POSIX shared memory) and page-locked in
    every process, so that each rank can download its share of the quotient polynomials over its own PCIe
    link straight to the final offsets (ms_commit_hooks.download_rank / download_world).  `array` is the
    uint8 view; rank 0 reads the proof from it after stark_prove_sharded returns."""

    def __init__(self, ctx, nbytes: int, dist, group=None):
        from multiprocessing import shared_memory

        self.ctx, self.dist, self.group = ctx, dist, group
        rank = dist.get_rank(group)
        names = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            names[0] = self.shm.name
            if os.environ.get("MINISTARK_SHM_INTERLEAVE", "0") != "0":
                _touch_interleaved(self.shm.buf, nbytes)  # spread the pages over the NUMA nodes before anyone pins them
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(names, src=src, group=group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=names[0])
            try:  # the creator unlinks; keep the resource tracker of the other ranks from doing it again
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = rank == 0
        self.nbytes = nbytes
        self.array = np.frombuffer(self.shm.buf, dtype=np.uint8, count=nbytes)
        ctx._check(ctx.lib.ms_host_register(ctx.h, C.c_void_p(self.array.ctypes.data), nbytes))
        self.registered = True
        dist.barrier(group=group)

    def close(self):
        if getattr(self, "registered", False):
            self.ctx.lib.ms_host_unregister(self.ctx.h, C.c_void_p(self.array.ctypes.data))
            self.registered = False
        self.dist.barrier(group=self.group)
        arr, self.array = self.array, None
        del arr
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.owner:
            self.shm.unlink()


def stark_prove_sharded(ctx, params, trace_cm, constraint_matrix: np.ndarray, out, dist=None, group=None,
                        proof_on_all_ranks: bool = False) -> int:
    """Stark::prove on this rank's replica with the commitments sharded over the process group.
    trace_cm: device [W, N] (every rank holds it).  out: a host uint8 array -- rank 0 then downloads the
    whole proof (the other ranks only with proof_on_all_ranks) -- or a SharedProofBuffer: every rank downloads
    1/world of the quotient polynomials into the one shared buffer in parallel (the proof bytes are ~all
    quotients, src/fri.rs:167, so the PCIe-bound tail of the proof shrinks by the number of GPUs).
    Returns the proof length."""
    w, n = trace_cm.shape
    m = np.ascontiguousarray(constraint_matrix, dtype=ctx.np_dtype).reshape(-1, w)
    ops = getattr(ctx, "_sharded_ops", None)
    if ops is None:
        ops = ctx._sharded_ops = CudaOps(ctx)
    com = ShardedCommitter(ops, int(params.trace_columns), int(params.inner_children) or 2, dist, group)
    com.validate(n, w, w + m.shape[0], int(params.blowup_factor))
    shared = isinstance(out, SharedProofBuffer)
    buf = out.array if shared else out
    if shared and com.world > 1:
        hooks = com.hooks(download=(com.rank, com.world))
    else:
        hooks = com.hooks(replica_only=(com.rank != 0 and not proof_on_all_ranks))
    cap = C.c_uint64(buf.size)
    rc = ctx.lib.ms_stark_prove_hooked(ctx.h, C.byref(params), C.c_void_p(trace_cm.data_ptr()), n, w, m.ctypes.data, m.shape[0],
                                       C.byref(hooks), buf.ctypes.data, C.byref(cap))
    if com.error is not None:
        raise com.error
    ctx._check(rc)
    if shared and com.world > 1:
        dist.barrier(group=group)  # every rank's share has landed in the shared buffer
    ctx.last_sharded_stats = dict(com.stats)
    return int(cap.value)
