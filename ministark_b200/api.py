"""Host-side handle on the C ABI: device memory comes from torch (plumbing only), every computation
is a call into libministark.so.  Matrices are torch tensors of shape [cols, rows] (column-major in the
reference's terms: one contiguous column per polynomial / register), dtype int64 (Goldilocks, the bit
pattern of the canonical uint64) or int32 (BabyBear)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import MiniStarkError, StarkParams

GOLDILOCKS, BABYBEAR = 0, 1
MODULUS = {GOLDILOCKS: 2**64 - 2**32 + 1, BABYBEAR: 2013265921}
EXT_DEGREE = {GOLDILOCKS: 2, BABYBEAR: 4}
_NP = {GOLDILOCKS: np.uint64, BABYBEAR: np.uint32}


def _torch():
    import torch

    return torch


class Context:
    """One prover context = one GPU + one stream (ms_ctx)."""

    def __init__(self, field: int, device: int = 0, use_torch_stream: bool = True):
        torch = _torch()
        if not torch.cuda.is_available():
            raise MiniStarkError(2, "no CUDA device: ministark_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.field = field
        self.device = device
        self.np_dtype = _NP[field]
        self.t_dtype = torch.int64 if field == GOLDILOCKS else torch.int32
        self.D = EXT_DEGREE[field]
        torch.cuda.set_device(device)
        stream = torch.cuda.current_stream(device).cuda_stream if use_torch_stream else None
        h = C.c_void_p()
        rc = self.lib.ms_ctx_create(field, device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise MiniStarkError(rc, "ms_ctx_create failed")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            ops = getattr(self, "_sharded_ops", None)
            if ops is not None:
                ops.close_peers()  # unmap / free the CUDA-IPC exchange buffers (all device work is done by now)
            # a communicator still bound here is dropped without the collective teardown (call comm_destroy on every
            # rank first for an orderly one)
            self.lib.ms_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- helpers
    def _check(self, rc: int):
        if rc != 0:
            raise MiniStarkError(rc, self.lib.ms_last_error(self.h).decode())

    def sync(self):
        self._check(self.lib.ms_sync(self.h))

    def launch_count(self) -> int:
        return int(self.lib.ms_launch_count(self.h))

    def set_profiling(self, on: bool):
        self._check(self.lib.ms_set_profiling(self.h, int(on)))

    def profile_collect(self):
        """{kernel name: (total device ms, launches)} since the last collect."""
        names = (C.c_char_p * 64)()
        ms_ = (C.c_float * 64)()
        cnt = (C.c_uint32 * 64)()
        n = self.lib.ms_profile_collect(self.h, names, ms_, cnt, 64)
        return {names[i].decode(): (float(ms_[i]), int(cnt[i])) for i in range(n)}

    def set_zero_display(self, empty: bool):
        self._check(self.lib.ms_set_zero_display(self.h, int(empty)))

    def set_transcript_option(self, option: int, value: int):
        """ms_set_transcript_option: 0-2 DigestBridge mask bytes, 3 leftover handling as published (1) / intended (0)"""
        self._check(self.lib.ms_set_transcript_option(self.h, option, value))

    def to_device(self, a: np.ndarray):
        torch = _torch()
        a = np.ascontiguousarray(a, dtype=self.np_dtype)
        signed = a.view(np.int64 if self.field == GOLDILOCKS else np.int32)
        return torch.from_numpy(signed).to(f"cuda:{self.device}")

    def to_host(self, t) -> np.ndarray:
        return t.detach().cpu().numpy().view(self.np_dtype)

    def empty(self, *shape):
        return _torch().empty(*shape, dtype=self.t_dtype, device=f"cuda:{self.device}")

    def zeros(self, *shape):
        return _torch().zeros(*shape, dtype=self.t_dtype, device=f"cuda:{self.device}")

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr())

    # ---------------------------------------------------------------- trace generation on the device (air.rs:73-112)
    def trace_synth(self, n: int, w: int, seed: int = 0x5EED000000000000):
        """the synthetic benchmark trace (synth.synth_trace) as a device [w, n] matrix, no upload"""
        out = self.empty(w, n)
        self._check(self.lib.ms_trace_synth(self.h, seed & (2**64 - 1), n, w, self._ptr(out)))
        return out

    def trace_recurrence(self, matrix, row0, steps: int, n: int, padding: int):
        """rows [0, steps): row_{i+1} = M row_i; rows [steps, n): padding.  Device [w, n]."""
        m = np.ascontiguousarray(matrix, dtype=self.np_dtype)
        w = m.shape[0]
        r0 = np.ascontiguousarray(row0, dtype=self.np_dtype)
        out = self.empty(w, n)
        self._check(self.lib.ms_trace_recurrence(self.h, m.ctypes.data, r0.ctypes.data, w, steps, n, int(padding), self._ptr(out)))
        return out

    # ---------------------------------------------------------------- stages
    def transpose_rm_to_cm(self, rm):
        """[rows, width] row-major -> [width, rows] (air.rs:151-153 gather)."""
        rows, width = rm.shape
        out = self.empty(width, rows)
        self._check(self.lib.ms_transpose_rm_to_cm(self.h, self._ptr(rm), rows, width, self._ptr(out)))
        return out

    def transpose_cm_to_rm(self, cm):
        width, rows = cm.shape
        out = self.empty(rows, width)
        self._check(self.lib.ms_transpose_cm_to_rm(self.h, self._ptr(cm), rows, width, self._ptr(out)))
        return out

    def merkle_commit(self, cm, leafs_per_node: int, inner_children: int = 2, deg: int = 1, want_nodes: bool = False):
        """MerkleTree::new (merkle.rs:81-148) over the row-major flattening of cm ([width*deg, rows])."""
        planes, rows = cm.shape
        width = planes // deg
        n_groups = rows * width // leafs_per_node if leafs_per_node else 0
        total = int(self.lib.ms_merkle_node_count(n_groups, inner_children)) if rows * width % max(leafs_per_node, 1) == 0 else 0
        nodes = None
        if want_nodes and total:
            nodes = _torch().empty(total, 8, dtype=_torch().int32, device=cm.device)
        root = (C.c_uint8 * 32)()
        self._check(self.lib.ms_merkle_commit(self.h, self._ptr(cm), cm.stride(0), rows, width, deg, leafs_per_node,
                                              inner_children, self._ptr(nodes) if nodes is not None else None, root))
        return (bytes(root), nodes) if want_nodes else bytes(root)

    @staticmethod
    def nodes_to_bytes(nodes) -> np.ndarray:
        """device digest words [n, 8] -> host bytes [n, 32] (big-endian words)."""
        w = nodes.detach().cpu().numpy().view(np.uint32)
        return w.astype(">u4").view(np.uint8).reshape(-1, 32)

    def intt_columns(self, evals_cm):
        cols, n = evals_cm.shape
        out = self.empty(cols, n)
        self._check(self.lib.ms_intt_columns(self.h, self._ptr(evals_cm), evals_cm.stride(0), n, cols, self._ptr(out), n))
        return out

    def coset_lde(self, coeffs_cm, blowup: int, shift: int, out=None):
        cols, n = coeffs_cm.shape
        if out is None:
            out = self.empty(cols, n * blowup)
        self._check(self.lib.ms_coset_lde(self.h, self._ptr(coeffs_cm), coeffs_cm.stride(0), n, cols, blowup, shift,
                                          self._ptr(out), out.stride(0)))
        return out

    def coset_lde_host(self, coeffs: np.ndarray, blowup: int, shift: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host-buffer entry point: poly-major [cols, n] in, row-major [L, cols] out (starks.rs:87-91)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=self.np_dtype)
        cols, n = coeffs.shape
        if out is None:
            out = np.empty((n * blowup, cols), dtype=self.np_dtype)
        self._check(self.lib.ms_coset_lde_host(self.h, coeffs.ctypes.data, n, cols, blowup, shift, out.ctypes.data))
        return out

    def linear_constraints(self, coeffs_cm, matrix: np.ndarray):
        """rows of `matrix` (T x W canonical scalars) -> T new coefficient columns (air.rs:130-134)."""
        w, n = coeffs_cm.shape
        m = np.ascontiguousarray(matrix, dtype=self.np_dtype)
        t = m.shape[0]
        assert m.shape[1] == w
        out = self.empty(t, n)
        self._check(self.lib.ms_linear_constraints(self.h, self._ptr(coeffs_cm), coeffs_cm.stride(0), n, w, m.ctypes.data, t,
                                                   self._ptr(out), n))
        return out

    def mix(self, coeffs_cm, r: int):
        cols, n = coeffs_cm.shape
        out = self.empty(n)
        self._check(self.lib.ms_mix(self.h, self._ptr(coeffs_cm), coeffs_cm.stride(0), n, cols, r, self._ptr(out)))
        return out

    def deep_open(self, coeffs_cm, z: np.ndarray) -> np.ndarray:
        """z: [Q, D] -> out [Q, cols, D] (starks.rs:140-151)."""
        cols, n = coeffs_cm.shape
        z = np.ascontiguousarray(z, dtype=self.np_dtype).reshape(-1, self.D)
        out = np.zeros((z.shape[0], cols, self.D), dtype=self.np_dtype)
        self._check(self.lib.ms_deep_open(self.h, self._ptr(coeffs_cm), coeffs_cm.stride(0), n, cols, z.ctypes.data, z.shape[0],
                                          out.ctypes.data))
        return out

    def fri_commit(self, poly_planes, domain: int, blowup: int):
        """poly_planes [D, domain/blowup] -> (codeword [D, domain], nodes [2*domain/2-1.., 8], root)."""
        torch = _torch()
        D, npad = poly_planes.shape
        assert D == self.D and npad * blowup == domain
        cw = self.empty(D, domain)
        total = int(self.lib.ms_merkle_node_count(domain // 2, 2))
        nodes = torch.empty(total, 8, dtype=torch.int32, device=poly_planes.device)
        root = (C.c_uint8 * 32)()
        self._check(self.lib.ms_fri_commit(self.h, self._ptr(poly_planes), poly_planes.stride(0), domain, blowup,
                                           self._ptr(cw), domain, self._ptr(nodes), root))
        return cw, nodes, bytes(root)

    def fri_deep_coeffs(self, poly_planes, z: Sequence[int]) -> np.ndarray:
        D, n = poly_planes.shape
        zz = np.ascontiguousarray(z, dtype=self.np_dtype)
        out = np.zeros((2, D), dtype=self.np_dtype)
        self._check(self.lib.ms_fri_deep_coeffs(self.h, self._ptr(poly_planes), poly_planes.stride(0), n, zz.ctypes.data,
                                                out.ctypes.data))
        return out

    def fri_fold(self, poly_planes, z, alpha, d):
        D, n = poly_planes.shape
        zz = np.ascontiguousarray(z, dtype=self.np_dtype)
        aa = np.ascontiguousarray(alpha, dtype=self.np_dtype)
        dd = np.ascontiguousarray(d, dtype=self.np_dtype)
        out = self.zeros(D, max(n // 2, 1))
        self._check(self.lib.ms_fri_fold(self.h, self._ptr(poly_planes), poly_planes.stride(0), n, zz.ctypes.data, aa.ctypes.data,
                                         dd.ctypes.data, self._ptr(out), out.stride(0)))
        return out

    def fri_query(self, prev_poly, prev_len: int, prev_cw, prev_nodes, next_cw, betas: Sequence[int], want_quot: bool = True):
        """One round of the query phase (fri.rs:132-176).  Returns dict(points [q,6,D], found [2q], neigh [2q,2,D],
        paths [2q, path_len, 2, 32] bytes, quot [q, prev_len-2, D])."""
        D, nd = prev_cw.shape
        q = len(betas)
        b = np.ascontiguousarray(betas, dtype=np.uint64)
        path_len = max((nd // 2).bit_length() - 1, 0)
        pts = np.zeros((q, 6, D), dtype=self.np_dtype)
        found = np.zeros(2 * q, dtype=np.uint64)
        neigh = np.zeros((2 * q, 2, D), dtype=self.np_dtype)
        paths = np.zeros((2 * q, path_len, 2, 32), dtype=np.uint8)
        nq = max(prev_len - 2, 0)
        quot = np.zeros((q, nq, D), dtype=self.np_dtype) if want_quot else None
        self._check(self.lib.ms_fri_query(self.h, self._ptr(prev_poly), prev_poly.stride(0), prev_len, self._ptr(prev_cw), prev_cw.stride(0),
                                          nd, self._ptr(prev_nodes), self._ptr(next_cw), next_cw.stride(0), b.ctypes.data, q,
                                          pts.ctypes.data, found.ctypes.data, neigh.ctypes.data, paths.ctypes.data,
                                          quot.ctypes.data if quot is not None and nq else None))
        return dict(points=pts, found=found, neigh=neigh, paths=paths, quot=quot)

    # ---------------------------------------------------------------- whole prover
    def stark_prove(self, params: StarkParams, trace_rm: np.ndarray, constraint_matrix: np.ndarray,
                    capacity: Optional[int] = None, constants: Optional[np.ndarray] = None) -> bytes:
        """constants: optional T additive constants of affine constraints (ms_stark_prove_affine)"""
        trace_rm = np.ascontiguousarray(trace_rm, dtype=self.np_dtype)
        n, w = trace_rm.shape
        m = np.ascontiguousarray(constraint_matrix, dtype=self.np_dtype).reshape(-1, w)
        cst = None if constants is None else np.ascontiguousarray(constants, dtype=self.np_dtype).reshape(m.shape[0])
        cap = C.c_uint64(capacity or (1 << 20))
        buf = np.empty(cap.value, dtype=np.uint8)

        def call():
            if cst is None:
                return self.lib.ms_stark_prove(self.h, C.byref(params), trace_rm.ctypes.data, n, w, m.ctypes.data, m.shape[0],
                                               buf.ctypes.data, C.byref(cap))
            return self.lib.ms_stark_prove_affine(self.h, C.byref(params), trace_rm.ctypes.data, n, w, m.ctypes.data, cst.ctypes.data,
                                                  m.shape[0], buf.ctypes.data, C.byref(cap))

        rc = call()
        if rc == 8:  # MS_ERR_BUFFER_TOO_SMALL: size written back
            buf = np.empty(cap.value, dtype=np.uint8)
            rc = call()
        self._check(rc)
        return buf[: cap.value].tobytes()

    def stark_prove_device(self, params: StarkParams, trace_cm, constraint_matrix: np.ndarray, out: np.ndarray) -> int:
        """trace_cm: device [W, N]; out: preallocated uint8 host buffer. Returns the proof length."""
        w, n = trace_cm.shape
        m = np.ascontiguousarray(constraint_matrix, dtype=self.np_dtype).reshape(-1, w)
        cap = C.c_uint64(out.size)
        self._check(self.lib.ms_stark_prove_device(self.h, C.byref(params), self._ptr(trace_cm), n, w, m.ctypes.data, m.shape[0],
                                                   out.ctypes.data, C.byref(cap)))
        return int(cap.value)

    def stark_verify(self, params: StarkParams, constrains_cm, proof: bytes, strict: bool = False) -> Tuple[bool, int]:
        """Stark::verify (starks.rs:171-235).  constrains_cm: device [C, N] coefficient columns (the Constrains object).
        Returns (accepted, reference line of the first failed check)."""
        cols, n = constrains_cm.shape
        raw = np.frombuffer(proof, dtype=np.uint8) if not isinstance(proof, np.ndarray) else proof
        ok, line = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.ms_stark_verify(self.h, C.byref(params), self._ptr(constrains_cm), constrains_cm.stride(0), n, cols,
                                             raw.ctypes.data, raw.size, int(strict), C.byref(ok), C.byref(line)))
        return bool(ok.value), int(line.value)

    # ---------------------------------------------------------------- multi-GPU (csrc/comm.cuh)
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL id made by rank 0; distribute it to the other ranks by any means."""
        buf = (C.c_uint8 * 128)()
        rc = _lib.load().ms_comm_unique_id(buf)
        if rc != 0:
            raise MiniStarkError(rc, "ms_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init_nccl(self, unique_id: bytes, rank: int, world: int):
        """one process per GPU: collective over all ranks"""
        buf = (C.c_uint8 * 128)(*unique_id)
        self._check(self.lib.ms_comm_init_nccl(self.h, buf, rank, world))

    @staticmethod
    def comm_init_local(ctxs: Sequence["Context"]):
        """one process, one host thread per rank: binds the contexts into a group (rank = position in the list)"""
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        rc = _lib.load().ms_comm_init_local(arr, len(ctxs))
        if rc != 0:
            raise MiniStarkError(rc, "ms_comm_init_local failed")

    def comm_destroy(self):
        self._check(self.lib.ms_comm_destroy(self.h))

    def comm_info(self) -> Tuple[int, int, str]:
        r, w, b = C.c_int32(), C.c_int32(), C.c_char_p()
        self.lib.ms_comm_info(self.h, C.byref(r), C.byref(w), C.byref(b))
        return r.value, w.value, b.value.decode()

    def set_shard_mask(self, mask: int):
        self._check(self.lib.ms_set_shard_mask(self.h, mask))

    def stark_prove_multi(self, params: StarkParams, trace_cm, constraint_matrix: np.ndarray, out: Optional[np.ndarray],
                          shared: bool = False, all_ranks: bool = False) -> int:
        """ms_stark_prove_multi on this rank (every rank of the communicator must call it).  out: host uint8 buffer
        (one buffer shared by all ranks with shared=True; may be None on ranks > 0 otherwise).  Returns the proof length."""
        w, n = trace_cm.shape
        m = np.ascontiguousarray(constraint_matrix, dtype=self.np_dtype).reshape(-1, w)
        cap = C.c_uint64(out.size if out is not None else 0)
        flags = (1 if shared else 0) | (2 if all_ranks else 0)
        self._check(self.lib.ms_stark_prove_multi(self.h, C.byref(params), self._ptr(trace_cm), n, w, m.ctypes.data, m.shape[0],
                                                  out.ctypes.data if out is not None else None, C.byref(cap), flags))
        return int(cap.value)

    def last_timings(self):
        names = (C.c_char_p * 64)()
        vals = (C.c_float * 64)()
        n = self.lib.ms_stark_last_timings(self.h, names, vals, 64)
        return [(names[i].decode(), float(vals[i])) for i in range(n)]
