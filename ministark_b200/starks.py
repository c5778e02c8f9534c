"""Mirror of the reference's `starks` module (src/starks.rs): `StarkConfig`, `Stark`, `StarkProof`.
`Stark.prove` hands the trace and the linear constraint matrix to ms_stark_prove; everything behind
that call runs on the GPU (plus the serial host transcript inside the library)."""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass
from typing import List, Optional, Tuple

from . import _lib
from ._lib import MiniStarkError, StarkParams
from .api import Context
from .field import FIELDS, StarkField

MAGIC = b"MSTARKP1"


@dataclass
class MerklePath:  # merkle.rs:293-298
    leaf_neighbours: List[tuple]
    path: List[List[bytes]]


@dataclass
class FriProof:  # fri.rs:18-22
    points: list
    queries: list
    quotients: list


@dataclass
class StarkProof:  # starks.rs:21-28
    arthur: bytes
    trace_commit: bytes
    constrain_trace_commit: bytes
    constrain_queries: List[List[tuple]]
    validity_queries: List[tuple]
    fri_proof: FriProof
    raw: bytes = b""

    @staticmethod
    def from_bytes(raw: bytes) -> "StarkProof":
        """Parse the canonical dump (DESIGN.md 'Proof bytes')."""
        assert raw[:8] == MAGIC
        pos = 8

        def take(n):
            nonlocal pos
            b = raw[pos : pos + n]
            assert len(b) == n
            pos += n
            return b

        rd64 = lambda: struct.unpack("<Q", take(8))[0]
        fid, D = struct.unpack("<II", take(8))
        F = FIELDS[fid]
        bs = F.base_bytes
        ext = lambda: tuple(int.from_bytes(take(bs), "little") for _ in range(D))
        arthur = take(rd64())
        tc, cc = take(32), take(32)
        Q, Cn = rd64(), rd64()
        assert Q * Cn * D * bs <= len(raw) - pos and (Cn > 0 or Q == 0)  # tampered counts: no allocation beyond the dump
        cq = [[ext() for _ in range(Cn)] for _ in range(Q)]
        vq = [ext() for _ in range(rd64())]
        points, queries, quotients = [], [], []
        for _ in range(rd64()):
            rp, rq, rquot = [], [], []
            for _ in range(rd64()):
                rp.append([(ext(), ext()) for _ in range(3)])
                paths = []
                for _ in range(2):
                    neigh = [ext() for _ in range(rd64())]
                    levels = [[take(32) for _ in range(rd64())] for _ in range(rd64())]
                    paths.append(MerklePath(neigh, levels))
                rq.append(paths)
                rquot.append([ext() for _ in range(rd64())])
            points.append(rp); queries.append(rq); quotients.append(rquot)
        assert pos == len(raw)
        return StarkProof(arthur, tc, cc, cq, vq, FriProof(points, queries, quotients), raw)


class StarkConfig:
    """StarkConfig::new(security_bits, blowup_factor, steps, trace_columns) (starks.rs:268-310).
    `inner_children` is an extension: the reference hard-wires 2 (starks.rs:283-302).  merkle.rs supports
    k-ary trees but only full ones (merkle.rs:93-104): BASELINE configs 3 and 5 as written (2^23 rows 4-ary,
    2^26 rows 8-ary) are rejected with its "Tree is not full!" (DESIGN.md section 8)."""

    def __init__(self, field: StarkField, security_bits: int, blowup_factor: int, steps: int, trace_columns: int,
                 inner_children: int = 2):
        self.field = field
        self.params = StarkParams(security_bits, blowup_factor, steps, trace_columns, inner_children)
        r, cq, fq = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = _lib.load().ms_stark_derive(field.field_id, C.byref(self.params), C.byref(r), C.byref(cq), C.byref(fq))
        if rc != 0:
            raise MiniStarkError(rc, "StarkConfig: bad parameters (security bits has to be at least 20)")
        self.rounds, self.constrain_queries, self.fri_queries = r.value, cq.value, fq.value
        self.degree = steps - 1

    @classmethod
    def new(cls, field, security_bits, blowup_factor, steps, trace_columns):
        return cls(field, security_bits, blowup_factor, steps, trace_columns)


class Stark:
    def __init__(self, config: StarkConfig, ctx: Optional[Context] = None):
        self.config = config
        self.ctx = ctx or Context(config.field.field_id)

    @classmethod
    def new(cls, config):
        return cls(config)

    def prove(self, air, witness) -> StarkProof:
        """Stark::prove (starks.rs:59-169)."""
        trace = air.trace(witness)
        matrix, consts = trace.affine_form()
        raw = self.ctx.stark_prove(self.config.params, trace.data, matrix, capacity=self.proof_bound(trace.length, trace.constrain_number()),
                                   constants=consts if consts.any() else None)
        return StarkProof.from_bytes(raw)

    def proof_bound(self, n: int, cols: int) -> int:
        return int(_lib.load().ms_stark_proof_bound(self.config.field.field_id, C.byref(self.config.params), n, cols))

    def verify(self, constrains, proof, strict: bool = False) -> bool:
        """Stark::verify (starks.rs:171-235): `constrains` is the Constrains object of `TraceTable.derive_constrains`,
        `proof` a StarkProof (its canonical dump is what crosses the C ABI).  Returns Ok(true) / raises on a failed
        assertion like the reference (its checks are `assert!`s).  strict also enforces the Merkle paths."""
        raw = proof.raw if isinstance(proof, StarkProof) else bytes(proof)
        ok, line = self.ctx.stark_verify(self.config.params, constrains.get_polynomials(), raw, strict=strict)
        if not ok:
            raise AssertionError(f"Stark::verify: check at reference line {line} failed" if line > 0 else "malformed proof dump")
        return True
