"""CPU: host-side logic that needs no GPU - the SamplingOp DAG planner, the packed index-set layout, the sampler
component's root sharding."""
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
