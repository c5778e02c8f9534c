"""The mirror of the reference's `merkle` and `util` modules (ministark_b200/merkle.py, util.py) against the reference's own unit
tests: src/util.rs:50-96 (exact values), src/merkle.rs:399-482 (node counts, parent-index arithmetic, proof round trip with
path lengths 3 and 2, panic for trees that are not full).

CPU half: the host side (index arithmetic, openings, check_proof) over digests made by the oracle.  GPU half (`-m gpu`):
`MerkleTree.new` computes the digests with ms_merkle_commit -- every node must equal the oracle's."""
import numpy as np
import pytest

from ministark_b200.field import BabyBear, Goldilocks
from ministark_b200.merkle import MerklePath, MerkleProofError, MerkleRoot, MerkleTree, MerkleTreeConfig, display
from ministark_b200.util import ceil_log2_k, is_power_of_two, logarithm_of_two_k

TWO, TWO_FOUR, FOUR, SIXTEEN = MerkleTreeConfig(2, 2), MerkleTreeConfig(4, 2), MerkleTreeConfig(4, 4), MerkleTreeConfig(16, 16)  # merkle.rs:349-375


def test_util_values_of_the_reference_unit_tests():
    assert all(is_power_of_two(x) for x in (0, 1, 2, 32, 128, 512, 1024))           # util.rs:51-58
    assert not is_power_of_two(24) and not is_power_of_two(48)                      # util.rs:60-61
    assert [logarithm_of_two_k(32, 2), logarithm_of_two_k(256, 4), logarithm_of_two_k(512, 8), logarithm_of_two_k(256, 16)] == [5, 4, 3, 2]
    for number, base, msg in [(6, 2, "number if not a power of 2"), (12, 4, "number if not a power of 2"), (32, 4, "number if not a power of base"),
                              (15, 8, "number if not a power of 2"), (16, 8, "number if not a power of base"),
                              (48, 16, "number if not a power of 2"), (64, 16, "number if not a power of base")]:   # util.rs:66-83
        with pytest.raises(ValueError, match=msg):
            logarithm_of_two_k(number, base)
    assert [ceil_log2_k(*a) for a in [(2, 2), (21, 2), (32, 2), (4, 4), (3, 4), (13, 4), (21, 4)]] == [1, 5, 5, 2, 2, 4, 6]  # util.rs:88-95
    with pytest.raises(AssertionError):
        ceil_log2_k(0, 2)
    with pytest.raises(AssertionError):
        logarithm_of_two_k(8, 3)


def _host_tree(pyref, F, leafs, cfg):
    """the mirror's host side over digests computed by the oracle (no GPU)"""
    o = pyref.MerkleTree(pyref.Goldilocks if F is Goldilocks else pyref.BabyBear, leafs, cfg.leafs_per_node, cfg.inner_children)
    return MerkleTree(list(leafs), list(o.nodes), cfg, o.levels), o


def test_display_equals_the_oracle(pyref):
    for F, RF in ((Goldilocks, pyref.Goldilocks), (BabyBear, pyref.BabyBear)):
        D = F.ext_degree
        for e in (0, 1, F.p - 1, 12345678901234567890 % F.p):
            assert display(e) == pyref.display(RF, e)
        for e in (tuple(range(D)), tuple([F.p - 1] * D), tuple([0] * D)):
            assert display(e) == pyref.display(RF, e)


def test_parent_index_arithmetic(pyref):
    """merkle.rs:421-461"""
    tree, _ = _host_tree(pyref, Goldilocks, list(range(16)), TWO)
    assert [tree.get_parent_idx(i) for i in (1, 4, 9, 13)] == [16, 18, 20, 22]
    assert [tree.get_parent_idx(i) for i in (16, 18, 20, 22)] == [24, 25, 26, 27]
    assert [tree.get_parent_idx(i) for i in (24, 25, 26, 27, 28, 29)] == [28, 28, 29, 29, 30, 30]
    tree, _ = _host_tree(pyref, Goldilocks, list(range(16)), TWO_FOUR)
    assert [tree.get_parent_idx(i) for i in (1, 4, 9, 13)] == [16, 17, 18, 19]
    assert [tree.get_parent_idx(i) for i in (16, 17, 18, 19, 20, 21)] == [20, 20, 21, 21, 22, 22]
    with pytest.raises(MerkleProofError, match="index outside of tree length"):
        tree.get_parent_idx(tree.get_node_number())
    with pytest.raises(MerkleProofError, match="index is root node"):
        tree.get_parent_idx(tree.get_node_number() - 1)


@pytest.mark.parametrize("cfg,nodes,total,path_len", [(TWO, 15, 31, 3), (TWO_FOUR, 7, 23, 2), (FOUR, 5, 21, 1), (SIXTEEN, 1, 17, 0)])
def test_node_counts_and_proof_round_trip(pyref, cfg, nodes, total, path_len):
    """merkle.rs:399-419 and :463-481 (path lengths 3 and 2 for TWO / TWO_FOUR)"""
    tree, o = _host_tree(pyref, Goldilocks, list(range(16)), cfg)
    assert (tree.get_node_number(), len(tree.leafs), len(tree.nodes)) == (total, 16, nodes)
    proof = tree.generate_proof(7)
    assert 7 in proof.leaf_neighbours and len(proof.path) == path_len
    assert MerkleRoot(tree.root()).check_proof(proof)
    ref = o.generate_proof(7)
    assert proof.leaf_neighbours == list(ref.leaf_neighbours) and proof.path == [list(l) for l in ref.path]
    # a wrong neighbour, a wrong sibling, a wrong root: rejected
    assert not MerkleRoot(tree.root()).check_proof(MerklePath([x + 1 for x in proof.leaf_neighbours], proof.path))
    if proof.path:
        bad = [list(l) for l in proof.path]
        bad[-1][0] = bytes(32)
        assert not MerkleRoot(tree.root()).check_proof(MerklePath(proof.leaf_neighbours, bad))
    assert not MerkleRoot(bytes(32)).check_proof(proof)
    with pytest.raises(MerkleProofError, match="leaf is not included in the tree"):
        tree.generate_proof(16)


def test_first_match_opens_duplicate_leaves(pyref):
    """get_leaf_index returns the FIRST equal leaf (merkle.rs:216-225): a duplicated value opens the earlier group"""
    leafs = [5, 6, 7, 8, 9, 6, 11, 12]
    tree, _ = _host_tree(pyref, Goldilocks, leafs, TWO)
    assert tree.get_leaf_index(6) == 1 and tree.generate_proof(6).leaf_neighbours == [5, 6]


def test_reference_quirk_parent_index_of_k_ary_trees_is_kept(pyref):
    """get_parent_idx's inner-node formula (merkle.rs:205-206) is only right for binary trees: in a 4-ary tree deeper than
    one inner level the path it walks is not the authentication path, and check_proof rejects the reference's own proof.
    The reference never meets this (StarkConfig::new builds binary trees, its (4,4) unit test has one inner level); the
    mirror keeps it, like every quirk of SURVEY.md App. B -- it must answer what the reference answers."""
    leafs = list(range(1000, 1000 + 1024))
    tree, o = _host_tree(pyref, Goldilocks, leafs, FOUR)
    assert tree.get_parent_idx(1024 + 194) == 1255 != 1024 + 256 + 194 // 4
    proof, ref = tree.generate_proof(leafs[777]), o.generate_proof(leafs[777])
    assert proof.path == [list(l) for l in ref.path]
    assert not MerkleRoot(tree.root()).check_proof(proof) and not pyref.check_proof(pyref.Goldilocks, o.root(), ref)
    # the same leaves under a binary inner tree: the round trip holds
    tree, _ = _host_tree(pyref, Goldilocks, leafs, TWO_FOUR)
    assert MerkleRoot(tree.root()).check_proof(tree.generate_proof(leafs[777]))


def test_extension_leaves(pyref):
    D = BabyBear.ext_degree
    leafs = [tuple((7 * i + d) % BabyBear.p for d in range(D)) for i in range(8)]
    tree, _ = _host_tree(pyref, BabyBear, leafs, TWO)
    proof = tree.generate_proof(leafs[5])
    assert proof.leaf_neighbours == leafs[4:6] and MerkleRoot(tree.root()).check_proof(proof)


def test_new_panics_for_trees_that_are_not_full():
    """merkle.rs:384-396: shape violations are found before anything touches the device"""
    with pytest.raises(AssertionError):
        MerkleTree.new(Goldilocks, [0, 1, 2], TWO)
    with pytest.raises(AssertionError, match="number if not a power of base"):
        MerkleTree.new(Goldilocks, list(range(32)), FOUR)   # 8 leaf groups are not a power of 4
    with pytest.raises(AssertionError):
        MerkleTree.new(Goldilocks, list(range(18)), FOUR)   # 18 leaves do not divide into groups of 4 (4 groups + 2)


def test_new_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ministark_b200 import MiniStarkError

    with pytest.raises(MiniStarkError):
        MerkleTree.new(Goldilocks, list(range(16)), TWO)


# ------------------------------------------------------------------------------------------ GPU half
@pytest.mark.gpu
@pytest.mark.parametrize("cfg,nodes,total,path_len", [(TWO, 15, 31, 3), (TWO_FOUR, 7, 23, 2), (FOUR, 5, 21, 1), (SIXTEEN, 1, 17, 0)])
def test_gpu_tree_runs_the_reference_unit_tests(pyref, cfg, nodes, total, path_len):
    """make_tree (merkle.rs:377-382) on the GPU: counts of test_node_calculation, test_check_proof, every digest equal to the oracle's"""
    tree = MerkleTree.new(Goldilocks, list(range(16)), cfg)
    o = pyref.MerkleTree(pyref.Goldilocks, list(range(16)), cfg.leafs_per_node, cfg.inner_children)
    assert tree.nodes == list(o.nodes) and tree.root() == o.root()
    assert (tree.get_node_number(), len(tree.leafs), len(tree.nodes)) == (total, 16, nodes)
    proof = tree.generate_proof(7)
    assert 7 in proof.leaf_neighbours and len(proof.path) == path_len
    assert MerkleRoot(tree.root()).check_proof(proof)


@pytest.mark.gpu
@pytest.mark.parametrize("field", [0, 1])
def test_gpu_tree_of_extension_leaves_and_edge_values(pyref, field):
    F, RF = (Goldilocks, pyref.Goldilocks) if field == 0 else (BabyBear, pyref.BabyBear)
    D = F.ext_degree
    rng = np.random.default_rng(5)
    edge = [0, 1, 9, 10, 99, 100, 10**9 - 1, 10**9, F.p - 1, F.p - 2]
    base = edge + [int(x) % F.p for x in rng.integers(0, 2**62, size=1024 - len(edge))]
    tree = MerkleTree.new(F, base, MerkleTreeConfig(4, 4))
    o = pyref.MerkleTree(RF, base, 4, 4)
    assert tree.nodes == list(o.nodes)
    assert tree.generate_proof(base[777]).path == [list(l) for l in o.generate_proof(base[777]).path]  # (quirk below included)
    tree = MerkleTree.new(F, base, MerkleTreeConfig(4, 2))
    assert tree.nodes == list(pyref.MerkleTree(RF, base, 4, 2).nodes)
    proof = tree.generate_proof(base[777])
    assert len(proof.path) == 8 and MerkleRoot(tree.root()).check_proof(proof)
    ext = [tuple(base[(D * i + d) % 1024] for d in range(D)) for i in range(256)]
    tree = MerkleTree.new(F, ext, TWO)          # the shape of a FRI round tree (starks.rs:290-295)
    o = pyref.MerkleTree(RF, ext, 2, 2)
    assert tree.nodes == list(o.nodes)
    proof = tree.generate_proof(ext[131])
    assert proof.leaf_neighbours == ext[130:132] and len(proof.path) == 7 and MerkleRoot(tree.root()).check_proof(proof)
