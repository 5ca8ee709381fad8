"""CPU: host-side logic that needs no GPU - the SamplingOp DAG planner, the packed index-set layout, the sampler
component's root sharding."""
import os

import numpy as np
import pytest

from gigl_b200 import dag
from gigl_b200.engine import unpack_tree

FOLLOWS, SHOWN, CLICKS = ("user", "follows", "user"), ("item", "shown_to", "user"), ("user", "clicks", "item")


def test_plan_orders_ops_and_numbers_calls():
    ops = [dag.SamplingOp("fof", FOLLOWS, 2, ["friends"]),          # listed before its input: the plan reorders
           dag.SamplingOp("friends", FOLLOWS, 3),
           dag.SamplingOp("seen", SHOWN, 2),
           dag.SamplingOp("their_clicks", CLICKS, 2, ["friends"], dag.OUTGOING)]
    planned = dag.plan(ops, "user")
    assert [p.key for p in planned] == ["friends", "seen", "their_clicks", "fof"]  # list order among the ops that are ready
    assert [p.call_no for p in planned] == [1, 2, 3, 4]
    assert planned[3].chain == ["friends", "fof"] and planned[3].fanouts == [3, 2] and planned[3].parent == "friends"
    assert planned[2].op.result_node_type == "item" and planned[2].op.frontier_node_type == "user"


def test_plan_expands_an_op_once_per_input():
    """GraphDBSampler.scala:66-82: an op with several input ops expands the union of their result nodes."""
    ops = [dag.SamplingOp("friends", FOLLOWS, 3),
           dag.SamplingOp("seen", SHOWN, 2),
           dag.SamplingOp("clickers", CLICKS, 2, ["seen"]),
           dag.SamplingOp("mixed", FOLLOWS, 2, ["friends", "clickers"]),
           dag.SamplingOp("deep", SHOWN, 4, ["mixed"])]
    planned = dag.plan(ops, "user")
    assert [p.key for p in planned] == ["friends", "seen", "clickers", "mixed@friends", "mixed@clickers", "deep@mixed@friends",
                                        "deep@mixed@clickers"]
    assert [p.call_no for p in planned] == [1, 2, 3, 4, 4, 5, 5]          # instances of one op share its call number
    by_key = {p.key: p for p in planned}
    assert by_key["deep@mixed@clickers"].chain == ["seen", "clickers", "mixed@clickers", "deep@mixed@clickers"]
    assert by_key["deep@mixed@clickers"].fanouts == [2, 2, 2, 4]
    res = {p.key: (np.zeros(1, np.int32), None, p.fanouts) for p in planned}
    enc = dag.encoder_ops(planned, res, {FOLLOWS: 0, SHOWN: 1, CLICKS: 2}, {"user": 0, "item": 1})
    assert [o["parent"] for o in enc] == [-1, -1, 1, 0, 2, 3, 4]
    assert [o["result_node_type"] for o in enc] == [0, 1, 0, 0, 0, 1, 1] and not any(o["outgoing"] for o in enc)


def test_plan_rejects_malformed_dags():
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", FOLLOWS, 2), dag.SamplingOp("a", FOLLOWS, 3)], "user")                 # duplicate names
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", FOLLOWS, 2), dag.SamplingOp("c", FOLLOWS, 2, ["a", "a"])], "user")     # the same input twice
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", FOLLOWS, 2, ["nope"])], "user")                                        # unknown input
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", SHOWN, 2), dag.SamplingOp("b", FOLLOWS, 2, ["a"])], "user")            # b expands users, a yields items
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", FOLLOWS, 2, ["b"]), dag.SamplingOp("b", FOLLOWS, 2, ["a"])], "user")   # cycle
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", FOLLOWS, 0)], "user")                                                  # fanout < 1
    with pytest.raises(ValueError):
        dag.plan([dag.SamplingOp("a", CLICKS, 2)], "user")                                                   # INCOMING over clicks expands items


def test_ops_from_config_reads_the_proto_yaml_form():
    path = {"rootNodeType": "user", "samplingOps": [
        {"opName": "a", "edgeType": {"srcNodeType": "user", "relation": "follows", "dstNodeType": "user"},
         "randomUniform": {"numNodesToSample": 5}},
        {"opName": "b", "edgeType": {"srcNodeType": "user", "relation": "clicks", "dstNodeType": "item"}, "inputOpNames": ["a"],
         "randomUniform": {"numNodesToSample": 2}, "samplingDirection": "OUTGOING"}]}
    ops = dag.ops_from_config(path)
    assert ops[0] == dag.SamplingOp("a", FOLLOWS, 5) and ops[1] == dag.SamplingOp("b", CLICKS, 2, ["a"], dag.OUTGOING)
    with pytest.raises(ValueError):
        dag.ops_from_config({"samplingOps": [{"opName": "t", "edgeType": path["samplingOps"][0]["edgeType"], "topK": {}}]})


@pytest.mark.parametrize("fan", [[4], [3, 2], [5, 1, 3]])
def test_unpack_tree_inverts_the_packed_layout(fan):
    """Packed form of gigl_infer_khop_sage_packed_host: per hop one count byte per parent slot + the filled slots in
    parent-slot order (a parent's filled slots are its first cnt slots)."""
    rng = np.random.default_rng(5)
    n_roots, width = 37, 1
    nbr, cnt, packed = [], [], []
    for f in fan:
        c = rng.integers(0, f + 1, n_roots * width).astype(np.int32)
        lvl = np.full((n_roots * width, f), -1, dtype=np.int32)
        for p, k in enumerate(c):
            lvl[p, :k] = rng.integers(0, 1000, k)
        packed.append(lvl[lvl >= 0])
        nbr.append(lvl.reshape(-1))
        cnt.append(c)
        width *= f
    got_nbr, got_cnt = unpack_tree(np.concatenate(packed), [c.astype(np.uint8) for c in cnt], fan)
    for h in range(len(fan)):
        assert np.array_equal(got_nbr[h], nbr[h]) and np.array_equal(got_cnt[h], cnt[h]) and got_cnt[h].dtype == np.int32
    with pytest.raises(ValueError):
        unpack_tree(np.concatenate(packed)[:-1], [c.astype(np.uint8) for c in cnt], fan)


@pytest.mark.parametrize("bits", [1, 7, 17, 22, 27, 31])
def test_unpack_bits_inverts_the_bit_stream(bits):
    """The bit stream of gigl_infer_khop_sage_bitpacked_host (entry e in bits [e * bits, (e + 1) * bits) of little-endian 32-bit
    words): the numpy unpacker and the C-ABI's host helper against a bit-by-bit packer."""
    import ctypes as C

    from gigl_b200 import _capi, unpack_bits

    rng = np.random.default_rng(bits)
    for n in (0, 1, 33, 1000):
        ids = rng.integers(0, 1 << bits, n, dtype=np.int64)
        stream = 0
        for e, v in enumerate(ids.tolist()):
            stream |= v << (e * bits)
        n_words = (n * bits + 31) // 32
        words = np.array([(stream >> (32 * j)) & 0xFFFFFFFF for j in range(n_words)], dtype=np.uint32)
        assert np.array_equal(unpack_bits(words, n, bits), ids.astype(np.int32))
        out = np.full(max(n, 1), -1, dtype=np.int32)
        buf = words if n_words else np.zeros(1, np.uint32)
        assert _capi.lib().gigl_unpack_bits_host(C.c_void_p(buf.ctypes.data), n, bits, C.c_void_p(out.ctypes.data)) == 0
        assert np.array_equal(out[:n], ids.astype(np.int32))
    assert _capi.lib().gigl_unpack_bits_host(None, 5, 0, None) != 0


def test_sampler_component_root_shares_partition_the_roots():
    from gigl_b200 import subgraph_sampler as ss

    ids = np.arange(0, 1003, dtype=np.int32) * 3
    try:
        for world in (1, 2, 3, 8):
            parts, quota = [], 0
            for rank in range(world):
                ss._SHARD = (rank, world)
                parts.append(ss._my_share(ids))
                quota += ss._my_quota(101)
                assert ss._my_quota(0) == 0
            assert np.array_equal(np.concatenate(parts), ids) and quota == 101
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    finally:
        ss._SHARD = (0, 1)


# ---- batch collation semantics pinned by the reference's own data-loader tests ---------------------------------------
# python/tests/unit/src/training/lib/data_loaders/rooted_node_neighborhood_batching_test.py builds three samples and
# asserts the node / edge counts of their collation; python/tests/unit/src/common/graph_builder/pyg_graph_builder_test.py
# asserts that adding the same graph twice changes nothing.  Here the same graphs in the sampler's tree form (an edge
# src -> dst hangs src under dst), fanouts [2, 1]:
#   triangle 0->1, 0->2, 1->2 : root 2, hop 1 = {0, 1}, hop 2 under 1 = {0}
#   line     3->4            : root 4, hop 1 = {3}
#   chain    1->2, 2->3      : root 3, hop 1 = {2}, hop 2 under 2 = {1}
_TREES = {"triangle": (2, [0, 1], [-1, 0]), "line": (4, [3, -1], [-1, -1]), "chain": (3, [2, -1], [1, -1])}


def _collate(names):
    from oracle import oracle as O

    roots = np.array([_TREES[n][0] for n in names], dtype=np.int32)
    nbr = [np.array(sum((_TREES[n][1] for n in names), []), dtype=np.int32), np.array(sum((_TREES[n][2] for n in names), []), dtype=np.int32)]
    node_ids, ei, root_idx = O.np_collate(roots, nbr, [2, 1])
    edges = sorted((int(node_ids[s]), int(node_ids[d])) for s, d in zip(ei[0], ei[1]))
    return node_ids, edges, root_idx, roots


def test_collation_counts_of_the_reference_batching_tests():
    # test_can_collate_correctly_without_edge_overlap: triangle + line -> 5 nodes, 4 edges
    node_ids, edges, root_idx, roots = _collate(["triangle", "line"])
    assert len(node_ids) == 5 and edges == [(0, 1), (0, 2), (1, 2), (3, 4)]
    assert np.array_equal(node_ids[root_idx], roots)
    # test_can_collate_correctly_with_edge_overlap: triangle + chain share 1->2 -> 4 nodes, 4 edges (the edge is not duplicated)
    node_ids, edges, root_idx, roots = _collate(["triangle", "chain"])
    assert sorted(node_ids.tolist()) == [0, 1, 2, 3] and edges == [(0, 1), (0, 2), (1, 2), (2, 3)]
    assert np.array_equal(node_ids[root_idx], roots)
    # pyg_graph_builder_test.test_can_create_with_preexisting_data_objects_filtering_existing_nodes_and_edges:
    # the same graph added twice is the graph itself
    once = _collate(["triangle"])
    twice = _collate(["triangle", "triangle"])
    assert sorted(once[0].tolist()) == sorted(twice[0].tolist()) == [0, 1, 2] and once[1] == twice[1] == [(0, 1), (0, 2), (1, 2)]
    assert len(np.unique(twice[0])) == len(twice[0])  # local ids are one per distinct node


def test_output_prefix_is_overwritten_not_appended_to(tmp_path):
    """TFRecordIO.scala:62 writes with mode("overwrite"): part files of an earlier run (another world size, batch size or
    sample limit) must not survive beside the new ones.  Every stale file has exactly one remover, so ranks need no
    collective: rank r its own files, rank 0 the un-ranked ones and those of ranks >= world."""
    from gigl_b200 import subgraph_sampler as S

    d = str(tmp_path / "samples")
    os.makedirs(d)
    stale = ["part-00000.tfrecord", "part-00007.tfrecord", "part-r000-00000.tfrecord", "part-r001-00003.tfrecord",
             "part-r005-00000.tfrecord", "_SUCCESS", "notes.txt"]

    def fill():
        for n in stale:
            open(os.path.join(d, n), "wb").write(b"x")

    old = S._SHARD
    try:
        fill()
        S._SHARD = (0, 1)
        S._prepare_dir(d)
        assert sorted(os.listdir(d)) == ["_SUCCESS", "notes.txt"]
        fill()
        S._SHARD = (1, 2)
        S._prepare_dir(d)  # rank 1 of 2 removes only its own
        assert "part-r001-00003.tfrecord" not in os.listdir(d) and "part-r000-00000.tfrecord" in os.listdir(d)
        assert "part-00000.tfrecord" in os.listdir(d) and "part-r005-00000.tfrecord" in os.listdir(d)
        S._SHARD = (0, 2)
        S._prepare_dir(d)  # rank 0 removes its own, the un-ranked files and those of ranks that no longer exist
        assert sorted(os.listdir(d)) == ["_SUCCESS", "notes.txt"]
    finally:
        S._SHARD = old
