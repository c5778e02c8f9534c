"""Mirror of the reference's `merkle` module (src/merkle.rs): `MerkleTreeConfig`, `MerkleTree` (`Tree::new / root /
get_node_number`, `generate_proof`), `MerklePath`, `MerkleRoot.check_proof`, `MerkleProofError`.

`MerkleTree.new` is the hot part (one SHA-256 over the decimal strings of every leaf group, then the inner levels,
merkle.rs:81-148): the digests come from ms_merkle_commit on the GPU, in the reference's node order (merkle.rs:119-140);
there is no CPU fallback.  Openings are host index arithmetic over those digests exactly as written in merkle.rs:188-288
(first leaf equal to the value, the leaf group, one sibling group per level); the prover inside the library does the same
on the device for the FRI queries (csrc/fri.cuh).  `check_proof` is the verifier's side (merkle.rs:312-338): hashlib."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np

from .api import Context
from .field import StarkField
from .util import logarithm_of_two_k

Elem = Union[int, tuple]


class MerkleProofError(Exception):  # error.rs:13-21
    def __init__(self, kind: str, msg: str):
        super().__init__(f"Error generating Merkle proof: {msg}")
        self.kind, self.msg = kind, msg


@dataclass(frozen=True)
class MerkleTreeConfig:  # merkle.rs:34-43
    leafs_per_node: int
    inner_children: int


def display(e: Elem) -> str:
    """`to_string()` of a field element (merkle.rs:165): ark's Display of Fp (canonical decimal) and of QuadExtField,
    nested for the Fp4 tower.  An extension element is the tuple of its coordinates in tower order."""
    if isinstance(e, (int, np.integer)):
        return str(int(e))
    if len(e) == 1:
        return str(int(e[0]))
    h = len(e) // 2
    return f"QuadExtField({display(tuple(e[:h]))} + {display(tuple(e[h:]))} * u)"


def calculate_from_leafs(children: Sequence[Elem]) -> bytes:  # merkle.rs:162-168
    h = hashlib.sha256()
    for c in children:
        h.update(display(c).encode())
    return h.digest()


def calculate_from_nodes(children: Sequence[bytes]) -> bytes:  # merkle.rs:171-177
    h = hashlib.sha256()
    for c in children:
        h.update(c)
    return h.digest()


@dataclass
class MerklePath:  # merkle.rs:293-298
    leaf_neighbours: List[Elem]
    path: List[List[bytes]]


class MerkleTree:
    """merkle.rs:56-66.  `leafs`: base-field integers, or tuples of D coordinates for extension elements."""

    def __init__(self, leafs: List[Elem], nodes: List[bytes], config: MerkleTreeConfig, levels: int):
        self.leafs, self.nodes, self.config, self.levels = leafs, nodes, config, levels

    @classmethod
    def new(cls, field: StarkField, inputs: Sequence[Elem], config: MerkleTreeConfig, ctx: Optional[Context] = None) -> "MerkleTree":
        """Tree::new (merkle.rs:81-148).  Panics (AssertionError) like the reference: a leaf count that is not a multiple of
        leafs_per_node, or a tree that is not full."""
        lpn, k = config.leafs_per_node, config.inner_children
        leaf_num = len(inputs)
        node_num = leaf_num // lpn
        try:
            levels = logarithm_of_two_k(node_num, k) + 1  # merkle.rs:93-96 (panics with the util.rs string)
        except ValueError as e:
            raise AssertionError(str(e)) from None
        assert leaf_num % lpn == 0                        # merkle.rs:99
        assert k ** (levels - 1) == leaf_num // lpn, f"Tree is not full! input length must be a power of {k}"  # merkle.rs:100-104
        leafs = [int(x) if isinstance(x, (int, np.integer)) else tuple(int(c) for c in x) for x in inputs]
        deg = 1 if isinstance(leafs[0], int) else len(leafs[0])
        own = ctx is None
        ctx = ctx or Context(field.field_id)
        try:
            planes = np.array([[x] if deg == 1 else list(x) for x in leafs], dtype=object).T  # [deg, leaf_num]
            cm = ctx.to_device(np.ascontiguousarray(planes.astype(ctx.np_dtype)))
            _root, dev_nodes = ctx.merkle_commit(cm, lpn, k, deg=deg, want_nodes=True)
            nodes = [bytes(row) for row in Context.nodes_to_bytes(dev_nodes)]  # (a device -> host copy: synchronises)
        finally:
            if own:
                ctx.close()
        assert len(nodes) == (1 - k ** levels) // (1 - k) and nodes[-1] == _root  # merkle.rs:116-118
        return cls(leafs, nodes, config, levels)

    def root(self) -> bytes:  # merkle.rs:151-154
        return self.nodes[-1]

    def get_node_number(self) -> int:  # merkle.rs:157-159
        return len(self.leafs) + len(self.nodes)

    def get_parent_idx(self, index: int) -> int:  # merkle.rs:188-209
        root_idx = self.get_node_number() - 1
        if index > root_idx:
            raise MerkleProofError("OutOfRangeError", "index outside of tree length")
        if index == root_idx:
            raise MerkleProofError("OutOfRangeError", "index is root node")
        if index < len(self.leafs):
            return len(self.leafs) + index // self.config.leafs_per_node
        return index + (self.get_node_number() - index + 1) // self.config.inner_children

    def get_leaf_index(self, node: Elem) -> int:  # merkle.rs:216-225 (first match)
        want = int(node) if isinstance(node, (int, np.integer)) else tuple(int(c) for c in node)
        for i, value in enumerate(self.leafs):
            if value == want:
                return i
        raise MerkleProofError("LeafNotFound", "leaf is not included in the tree")

    def get_leaf_neighbours(self, index: int) -> List[Elem]:  # merkle.rs:230-236
        n = self.config.leafs_per_node
        start = index - index % n
        return list(self.leafs[start:start + n])

    def get_inner_neighbours(self, index: int) -> List[bytes]:  # merkle.rs:241-248
        shifted = index - len(self.leafs)
        n = self.config.inner_children
        start = shifted - shifted % n
        return list(self.nodes[start:start + n])

    def calculate_path(self, index: int) -> List[List[bytes]]:  # merkle.rs:253-265
        path, current = [], index
        for _ in range(1, self.levels):
            path.append(self.get_inner_neighbours(current))
            current = self.get_parent_idx(current)
        return path

    def generate_proof(self, leaf: Elem) -> MerklePath:  # merkle.rs:272-288
        leaf_index = self.get_leaf_index(leaf)
        leaf_neighbours = self.get_leaf_neighbours(leaf_index)
        leaf_parent = self.get_parent_idx(leaf_index)
        return MerklePath(leaf_neighbours, self.calculate_path(leaf_parent))


@dataclass
class MerkleRoot:  # merkle.rs:302
    hash: bytes

    def check_proof(self, proof: MerklePath) -> bool:  # merkle.rs:312-338
        previous = calculate_from_leafs(proof.leaf_neighbours)
        for level in proof.path:
            if previous not in level:
                return False
            previous = calculate_from_nodes(level)
        return previous == self.hash
